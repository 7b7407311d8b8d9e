"""ctypes binding of libkobato_b200.so (the C ABI declared in include/kobato_b200.h).

There is NO fallback: if the shared library is missing or no sm_100 device is present, every
entry point raises ``KobatoNativeError``.  PyTorch is used only by callers for tensor handoff
(``tensor.data_ptr()`` and the current stream handle); this module itself does not import torch.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

KE_OK, KE_E_INVALID, KE_E_CUDA, KE_E_CAPACITY, KE_E_NOMEM, KE_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
KE_JOIN_REQUIRE_BAND = 1
KE_OPT_PHASH_GENERIC = 1
KE_OPT_JOIN_MODE = 2
KE_OPT_SSIM_V1 = 3
KE_OPT_RESIZE_GENERIC = 4
KE_OPT_PHASH_CFG = 5
KE_ABI_VERSION = 2

import os as _os

# KE_LIB_PATH: an alternative build of the same library (A/B timing of kernel variants from tools/); never a fallback
_LIB_PATH = Path(_os.environ.get("KE_LIB_PATH") or Path(__file__).resolve().parent / "libkobato_b200.so")
_lib = None
_lock = threading.RLock()  # re-entrant: group() builds Context objects (which call load()) while holding it


class KobatoNativeError(RuntimeError):
    """The CUDA library is missing, failed to load, or a call returned an error status."""

    def __init__(self, message: str, status: int = KE_E_CUDA):
        super().__init__(message)
        self.status = status


class CapacityError(KobatoNativeError):
    """An output buffer was too small; ``required`` holds the count that would have fit."""

    def __init__(self, message: str, required: int):
        super().__init__(message, KE_E_CAPACITY)
        self.required = required


_SIGNATURES = {
    "ke_abi_version": (C.c_int, []),
    "ke_last_error": (C.c_char_p, []),
    "ke_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ke_ctx_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "ke_ctx_device_count": (C.c_int, [C.c_void_p]),
    "ke_ctx_child": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ke_ctx_destroy": (None, [C.c_void_p]),
    "ke_ctx_device": (C.c_int, [C.c_void_p]),
    "ke_ctx_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ke_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "ke_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "ke_resample_ksize": (C.c_int, [C.c_int, C.c_int]),
    "ke_resample_table": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "ke_phash_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ke_phash_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "ke_hamming_join": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_void_p]),
    "ke_hamming_join_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_int64)]),
    "ke_hamming_join_pairs": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "ke_ssim_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "ke_ssim_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "ke_luma_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_void_p]),
    "ke_orb_match_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ke_cluster_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ke_scan_table_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ke_gray_resize_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ke_tile_ahash_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ke_bits_hamming_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "ke_plane_sad_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p]),
    "ke_cluster_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "ke_synth_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                  C.c_int64, C.c_uint64, C.c_int, C.c_void_p]),
    "ke_microbench_popc": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

EXPORTS = tuple(_SIGNATURES)


def library_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    with _lock:
        if _lib is None:
            if not _LIB_PATH.exists():
                raise KobatoNativeError(
                    f"{_LIB_PATH} is missing: build it with `python kobato-eyes_b200/csrc/build.py` "
                    "(there is no CPU fallback)")
            try:
                lib = C.CDLL(str(_LIB_PATH))
            except OSError as exc:  # pragma: no cover - depends on the box
                raise KobatoNativeError(f"cannot load {_LIB_PATH}: {exc}") from exc
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.ke_abi_version() != KE_ABI_VERSION:
                raise KobatoNativeError("libkobato_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def last_error() -> str:
    msg = load().ke_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> None:
    if status == KE_OK:
        return
    msg = last_error() or what
    if status == KE_E_INVALID:
        raise ValueError(msg)
    raise KobatoNativeError(f"{what}: {msg} (status {status})", status)


class ScanStats(C.Structure):
    """``ke_scan_stats`` of include/kobato_b200.h."""

    _fields_ = [(name, C.c_int64) for name in ("n_buckets", "buckets_ge2", "max_bucket", "candidates", "after_same_id",
                                               "edges", "members", "clusters")]

    def as_dict(self) -> dict:
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


class Context:
    """Owns one ``ke_ctx``: one device, or several (``ke_ctx_create_multi``: the ``_host`` entry points then fan their
    units over all of them, ``child(k)`` drives device k with ``d_`` pointers)."""

    def __init__(self, device=0, *, _borrowed=None):
        lib = load()
        if _borrowed is not None:  # a child of a multi-device context: not ours to destroy
            self._h, self._owned = C.c_void_p(_borrowed), False
        else:
            devices = [int(device)] if isinstance(device, int) else [int(d) for d in device]
            arr = (C.c_int * len(devices))(*devices)
            handle = C.c_void_p()
            check(lib.ke_ctx_create_multi(arr, len(devices), C.byref(handle)), "ke_ctx_create_multi")
            self._h, self._owned = handle, True
        self.device = int(lib.ke_ctx_device(self._h))
        self.n_devices = int(lib.ke_ctx_device_count(self._h))
        self.sm_count = lib.ke_ctx_sm_count(self._h)
        self.lock = threading.Lock()

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise KobatoNativeError("context already destroyed")
        return self._h

    @property
    def devices(self) -> list[int]:
        lib = load()
        return [int(lib.ke_ctx_device(C.c_void_p(lib.ke_ctx_child(self.handle, k)))) for k in range(self.n_devices)]

    def child(self, k: int) -> "Context":
        h = load().ke_ctx_child(self.handle, int(k))
        if not h:
            raise IndexError(k)
        return self if k == 0 else Context(_borrowed=h)

    @property
    def launches(self) -> int:
        return int(load().ke_ctx_launch_count(self.handle))

    def set_option(self, option: int, value: int) -> None:
        check(load().ke_ctx_set_option(self.handle, int(option), int(value)), "ke_ctx_set_option")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            if self._owned:
                load().ke_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_contexts: dict[int, Context] = {}
_group: Context | None = None


def _default_devices() -> list[int]:
    """Devices of the process-wide multi-device context: ``KE_DEVICES=0,2,3`` if set; under a one-process-per-GPU
    launcher (``WORLD_SIZE`` > 1: bench.py, pipeline.scan) this rank's device only; otherwise every visible device —
    the reference calls this path from one worker thread of one process (src/ui/dup_tab.py:118)."""
    import os
    import sys

    env = os.environ.get("KE_DEVICES", "").strip()
    if env and env != "all":
        return [int(x) for x in env.split(",") if x.strip()]
    torch = sys.modules.get("torch")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not env:
        if torch is not None and torch.cuda.is_available():
            return [torch.cuda.current_device()]
        return [int(os.environ.get("LOCAL_RANK", "0"))]
    if torch is not None and torch.cuda.is_available():
        return list(range(torch.cuda.device_count()))
    try:
        import ctypes.util

        rt = C.CDLL(ctypes.util.find_library("cudart") or "libcudart.so")
        n = C.c_int(0)
        if rt.cudaGetDeviceCount(C.byref(n)) == 0 and n.value > 0:
            return list(range(n.value))
    except OSError:
        pass
    return [0]


def group(devices=None) -> Context:
    """The process-wide multi-device context behind the host-array entry points (``ops.*`` on numpy inputs and the
    drop-in modules on top of them).  ``devices`` (first call only) overrides ``KE_DEVICES`` / the default."""
    global _group
    with _lock:
        g = _group
    if g is None:
        g = Context(list(devices) if devices is not None else _default_devices())
        with _lock:
            if _group is None:
                _group = g
                for k, dev in enumerate(g.devices):  # device tensors on these GPUs share the children
                    _contexts.setdefault(dev, g.child(k))
            g = _group
    return g


def reset_group(devices=None) -> Context:
    """Drop the process-wide context (tests) and build a new one over ``devices``."""
    global _group
    with _lock:
        old, _group = _group, None
        if old is not None:
            for dev in list(_contexts):
                if _contexts[dev] is old or not _contexts[dev]._owned:
                    del _contexts[dev]
    if old is not None:
        old.close()
    return group(devices)


def context(device: int | None = None) -> Context:
    """Process-wide single-device context for ``device`` (default: torch's current device if torch is loaded, else 0)."""
    if device is None:
        device = 0
        import sys

        torch = sys.modules.get("torch")
        if torch is not None and torch.cuda.is_available():
            device = torch.cuda.current_device()
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts.setdefault(device, ctx)
            ctx = _contexts[device]
    return ctx
