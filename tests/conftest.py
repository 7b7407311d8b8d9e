"""Shared test plumbing: import paths, the `gpu` marker, golden-vector loaders."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "kobato-eyes_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = Path(__file__).resolve().parent / "golden"
REFERENCE_SRC = Path("/root/reference/src")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the live reference under /root/reference")


def pytest_collection_modifyitems(config, items):
    have_gpu = False
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not mounted")
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not REFERENCE_SRC.exists():
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_phash():
    return json.loads((GOLDEN / "phash_golden.json").read_text())


@pytest.fixture(scope="session")
def golden_planes():
    return np.load(GOLDEN / "planes_golden.npz")


@pytest.fixture(scope="session")
def golden_scanner():
    return json.loads((GOLDEN / "scanner_golden.json").read_text())


@pytest.fixture(scope="session")
def golden_ssim():
    return json.loads((GOLDEN / "ssim_golden.json").read_text())


@pytest.fixture(scope="session")
def reference_modules():
    """The live reference's hot-path modules (build container only)."""
    if not REFERENCE_SRC.exists():
        pytest.skip("/root/reference not mounted")
    sys.path.append(str(REFERENCE_SRC))
    import importlib

    return {
        "phash": importlib.import_module("sig.phash"),
        "scanner": importlib.import_module("dup.scanner"),
        "fastsig": importlib.import_module("core.fastsig"),
    }


def normalise_clusters(clusters):
    """Order-insensitive form: {keeper: [(file_id, best_hamming) in listed order]} plus the
    multiset of member sets.  The reference's cluster list order has set-iteration-dependent
    ties (src/dup/scanner.py:315-318), so lists are compared modulo cluster order."""
    return sorted((c["keeper"], tuple(tuple(m) for m in c["members"])) for c in clusters)
