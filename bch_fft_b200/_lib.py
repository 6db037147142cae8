"""ctypes binding of the C-ABI CUDA library (include/bchfft_cuda.h).

The product loads exactly one file: bch_fft_b200/csrc/libbchfft_cuda.so, built by nvcc for
sm_100a (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
bch_fft_b200/csrc``).  If it is missing, or no CUDA device is visible, calls fail loudly --
there is no CPU fallback.  (The unit tests may bind the CPU-emulator build of the same
sources through ``bind(path)``; nothing in this package does.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_SO = os.path.join(HERE, "csrc", "libbchfft_cuda.so")

_vp = C.c_void_p
_sz = C.c_size_t
_u32 = C.c_uint32

_SIGNATURES = {
    "bchfft_last_error": (C.c_char_p, []),
    "bchfft_set_last_error": (None, [C.c_char_p]),
    "bchfft_device_count": (C.c_int, []),
    "bchfft_is_emulated": (C.c_int, []),
    "bchfft_field_create": (C.c_int, [C.c_int, _u32, _vp, _vp, C.POINTER(_vp)]),
    "bchfft_field_destroy": (None, [_vp]),
    "bchfft_field_sync": (C.c_int, [_vp]),
    "bchfft_additive_fft_host": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _u32, _sz]),
    "bchfft_additive_fft_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _u32, _sz, _vp]),
    "bchfft_multiplicative_fft_host": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _u32, _sz]),
    "bchfft_multiplicative_fft_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _u32, _sz, _vp]),
    "bchfft_code_create": (C.c_int, [_vp, _u32, _u32, _vp, _u32, C.POINTER(_vp)]),
    "bchfft_code_destroy": (None, [_vp]),
    "bchfft_bch_decode_host": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "bchfft_bch_decode_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "bchfft_bch_encode_host": (C.c_int, [_vp, _vp, _sz, _u32, _vp, _sz]),
    "bchfft_bch_encode_dev": (C.c_int, [_vp, _vp, _sz, _u32, _vp, _sz, _vp]),
    "bchfft_bch_syndrome_host": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "bchfft_bch_elp_host": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "bchfft_bch_locate_host": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "bchfft_dev_alloc": (C.c_int, [C.c_int, _sz, C.POINTER(_vp)]),
    "bchfft_dev_free": (C.c_int, [C.c_int, _vp]),
    "bchfft_pinned_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "bchfft_pinned_free": (C.c_int, [_vp]),
    "bchfft_set_concurrent_hosts": (None, [C.c_int]),
    "bchfft_dev_upload": (C.c_int, [C.c_int, _vp, _vp, _sz]),
    "bchfft_dev_download": (C.c_int, [C.c_int, _vp, _vp, _sz]),
    "bchfft_profile_begin": (C.c_int, []),
    "bchfft_profile_end": (C.c_int, []),
    "bchfft_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                     C.POINTER(C.c_uint64)]),
    "bchfft_launch_count": (C.c_uint64, []),
    "bchfft_launch_count_reset": (None, []),
}

EXPORTS = tuple(_SIGNATURES)


class BchFftError(RuntimeError):
    pass


def bind(path: str) -> C.CDLL:
    """dlopen `path` and attach the C-ABI prototypes; raises if any declared symbol is missing."""
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(CUDA_SO):
            raise BchFftError(
                f"{CUDA_SO} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  bch_fft_b200 has no CPU fallback.")
        _lib = bind(CUDA_SO)
    return _lib


def profile(l: C.CDLL, fn):
    """run fn() with per-kernel timing; returns {kernel: (total_ms, launches)}"""
    l.bchfft_profile_begin()
    try:
        fn()
    finally:
        n = l.bchfft_profile_end()
    out = {}
    for i in range(max(n, 0)):
        name, ms, cnt = C.c_char_p(), C.c_double(), C.c_uint64()
        l.bchfft_profile_get(i, C.byref(name), C.byref(ms), C.byref(cnt))
        out[name.value.decode()] = (ms.value, cnt.value)
    return out


def check(l: C.CDLL, rc: int, what: str) -> None:
    if rc != 0:
        raise BchFftError(f"{what} failed ({rc}): {l.bchfft_last_error().decode()}")
