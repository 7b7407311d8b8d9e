"""kobato_b200 — B200 (sm_100a) implementation of kobato-eyes' duplicate-detection hot path.

Layout mirrors the reference's packages for the path (reference paths under src/):

    kobato_b200.sig.phash      <- sig/phash.py        phash, dhash, hamming64
    kobato_b200.core.fastsig   <- core/fastsig.py     compute_signatures_mp, fast_fill_missing_signatures, ...
    kobato_b200.core.signature <- core/signature.py   compute_signatures_from_image, ensure_signatures
    kobato_b200.dup.scanner    <- dup/scanner.py      DuplicateFile, DuplicateScanConfig, DuplicateScanner, ...
    kobato_b200.dup.refine     <- dup/refine.py       RefinementThresholds, RefinedMatch, refine_pair (+ refine_pairs_batch)
    kobato_b200.dup.cluster    <- dup/cluster.py      Cluster, ClusterBuilder

``kobato_b200.ops`` holds the array-level calls, ``kobato_b200._native`` the ctypes binding of
``libkobato_b200.so``.  Everything numeric runs in hand-written CUDA kernels; there is no CPU
fallback (missing library or device => ``KobatoNativeError``).
"""
from ._native import CapacityError, KobatoNativeError  # noqa: F401

__all__ = ["KobatoNativeError", "CapacityError"]
__version__ = "0.1.0"
