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

_LIB_PATH = Path(__file__).resolve().parent / "libkobato_b200.so"
_lib = None
_lock = threading.Lock()


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
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ke_ssim_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "ke_gray_resize_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ke_tile_ahash_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ke_bits_hamming_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "ke_plane_sad_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p]),
    "ke_cluster_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "ke_synth_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64,
                                  C.c_uint64, C.c_int, C.c_void_p]),
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
            if lib.ke_abi_version() != 1:
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


class Context:
    """Owns one ``ke_ctx`` (one per process and device)."""

    def __init__(self, device: int = 0):
        lib = load()
        handle = C.c_void_p()
        check(lib.ke_ctx_create(int(device), C.byref(handle)), "ke_ctx_create")
        self._h = handle
        self.device = int(device)
        self.sm_count = lib.ke_ctx_sm_count(handle)
        self.lock = threading.Lock()

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise KobatoNativeError("context already destroyed")
        return self._h

    @property
    def launches(self) -> int:
        return int(load().ke_ctx_launch_count(self.handle))

    def set_option(self, option: int, value: int) -> None:
        check(load().ke_ctx_set_option(self.handle, int(option), int(value)), "ke_ctx_set_option")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            load().ke_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def context(device: int | None = None) -> Context:
    """Process-wide context for ``device`` (default: torch's current device if torch is loaded, else 0)."""
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
