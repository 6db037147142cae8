"""Python mirror of the reference's C interface (include/finite_field.h, additive_fft.h,
multiplicative_fft.h, bch.h) plus the batch / device-resident entry points.

Two layers are exposed, both thin ctypes bindings -- Python does no arithmetic here:

* ``Host``: host/libbchfft_host.so, the drop-in C library with the reference's own symbols
  (``evaluate_polynomial_additive_FFT``, ``bch_decode``, ...) and the batch extensions;
  it marshals to the CUDA layer.
* ``DeviceField`` / ``DeviceCode``: the C-ABI of include/bchfft_cuda.h directly, for callers that
  keep data resident in HBM (bench.py passes torch device pointers and torch's stream).

Function names, argument meaning and error behaviour follow the reference
(libfft/additive_fft.c:13-21, libbch/bch.c:30-54, 492-562).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(os.path.dirname(HERE), "host", "libbchfft_host.so")

u32p = C.POINTER(C.c_uint32)
ulp = C.POINTER(C.c_ulong)


class TfiniteField(C.Structure):
    """include/finite_field.h:13-26"""
    _fields_ = [("m", C.c_uint32), ("n", C.c_uint32), ("logzero", C.c_uint32),
                ("exp", u32p), ("log", u32p)]


class TbchCode(C.Structure):
    """include/bch.h:10-22"""
    _fields_ = [("m_field", TfiniteField), ("m_distance", C.c_uint32),
                ("m_packed_gen_poly", ulp), ("m_gen_poly_degree", C.c_uint32),
                ("m_length", C.c_uint32)]


class TbchCodeBuffers(C.Structure):
    """include/bch.h:25-36"""
    _fields_ = [("m_syndrome", u32p), ("m_buffer_1", u32p), ("m_buffer_2", u32p),
                ("m_A", u32p), ("m_B", u32p), ("m_C_elp", u32p), ("m_error", u32p)]


def _u32(a):
    a = np.ascontiguousarray(a, np.uint32)
    return a, a.ctypes.data_as(u32p)


class Host:
    """Binding of the drop-in host library.  ``Host()`` opens the real one (CUDA behind it);
    tests may pass the path of the CPU-emulator build."""

    def __init__(self, path: str | None = None):
        path = path or HOST_SO
        if not os.path.exists(path):
            raise _lib.BchFftError(
                f"{path} is missing: run `make` (or __graft_entry__.build()).  There is no "
                "CPU fallback.")
        L = self.lib = C.CDLL(path)
        fp = C.POINTER(TfiniteField)
        cp = C.POINTER(TbchCode)
        bp = C.POINTER(TbchCodeBuffers)
        L.create_binary_finite_field.restype = TfiniteField
        L.create_binary_finite_field.argtypes = [C.c_uint32]
        L.free_finite_field.argtypes = [fp]
        L.multiply.restype = C.c_uint32
        L.multiply.argtypes = [fp, C.c_uint32, C.c_uint32]
        L.evaluate_polynomial_additive_FFT.restype = None
        L.evaluate_polynomial_additive_FFT.argtypes = [fp, u32p, u32p, C.c_uint32]
        for name in ("additive_DFT_opt", "additive_DFT", "additive_DFT_ref"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [fp, u32p, C.c_uint32]
        L.additive_fft_batch.restype = C.c_int
        L.additive_fft_batch.argtypes = [fp, u32p, u32p, C.c_uint32, C.c_size_t]
        L.multiplicative_FFT.restype = C.c_int
        L.multiplicative_FFT.argtypes = [fp, u32p, u32p]
        L.evaluate_polynomial_multiplicative_FFT.restype = C.c_int
        L.evaluate_polynomial_multiplicative_FFT.argtypes = [fp, u32p, u32p, C.c_uint32, u32p]
        L.multiplicative_fft_batch.restype = C.c_int
        L.multiplicative_fft_batch.argtypes = [fp, u32p, u32p, C.c_uint32, C.c_size_t]
        L.bch_gen_code.restype = TbchCode
        L.bch_gen_code.argtypes = [C.c_uint32, C.c_uint32]
        L.bch_free_code.argtypes = [cp]
        L.bch_encode.restype = C.c_int
        L.bch_encode.argtypes = [cp, ulp, u32p, C.c_uint32]
        L.bch_alloc_buffers.restype = TbchCodeBuffers
        L.bch_alloc_buffers.argtypes = [cp]
        L.bch_free_buffers.argtypes = [bp]
        L.bch_eval_syndrome.restype = C.c_int32
        L.bch_eval_syndrome.argtypes = [cp, bp, ulp]
        L.bch_compute_elp.restype = C.c_uint32
        L.bch_compute_elp.argtypes = [cp, bp]
        L.bch_locate_errors.restype = C.c_uint32
        L.bch_locate_errors.argtypes = [cp, bp, C.c_uint32]
        L.bch_decode.restype = C.c_uint32
        L.bch_decode.argtypes = [cp, ulp, bp, u32p]
        L.bch_decode_batch.restype = C.c_int
        L.bch_decode_batch.argtypes = [cp, ulp, C.c_size_t, C.POINTER(C.c_ubyte), u32p]
        L.bch_encode_batch.restype = C.c_int
        L.bch_encode_batch.argtypes = [cp, ulp, ulp, C.c_size_t, C.c_uint32, C.c_size_t]
        L.bchfft_last_error.restype = C.c_char_p
        L.bchfft_host_shutdown.restype = None
        L.bchfft_host_num_devices.restype = C.c_int
        L.bchfft_host_num_contexts.restype = C.c_int

    def _check(self, rc, what):
        if rc:
            raise _lib.BchFftError(f"{what} failed ({rc}): {self.lib.bchfft_last_error().decode()}")

    # ---- fields -------------------------------------------------------------------------
    def create_binary_finite_field(self, m: int) -> "FiniteField":
        return FiniteField(self, m)

    # ---- codes --------------------------------------------------------------------------
    def bch_gen_code(self, codeword_size: int, minimum_distance: int) -> "BchCode":
        return BchCode(self, codeword_size, minimum_distance)

    def shutdown(self):
        self.lib.bchfft_host_shutdown()


class FiniteField:
    """create_binary_finite_field (libbase/finite_field.c:5-49)"""

    def __init__(self, host: Host, m: int):
        self.host = host
        self.c = host.lib.create_binary_finite_field(m)
        if not self.c.exp:
            raise ValueError(f"no reduction polynomial for GF(2^{m})")
        self.m, self.n = self.c.m, self.c.n
        self.exp = np.ctypeslib.as_array(self.c.exp, (self.n,))
        self.log = np.ctypeslib.as_array(self.c.log, (self.n,))

    def multiply(self, a, b):
        return self.host.lib.multiply(C.byref(self.c), a, b)

    def evaluate_polynomial_additive_FFT(self, poly: np.ndarray, num_terms: int) -> np.ndarray:
        """libfft/additive_fft.c:13-21: returns po_result (n-1 entries); `poly` (uint32[n],
        C-contiguous) is clobbered with the natural-order values, like the reference."""
        assert poly.dtype == np.uint32 and poly.size == self.n and poly.flags["C_CONTIGUOUS"]
        out = np.zeros(self.n, np.uint32)
        self.host.lib.evaluate_polynomial_additive_FFT(C.byref(self.c), poly.ctypes.data_as(u32p),
                                                       out.ctypes.data_as(u32p), num_terms)
        return out[: self.n - 1]

    def additive_DFT(self, poly: np.ndarray, degree: int, variant: str = "additive_DFT_opt"):
        assert poly.dtype == np.uint32 and poly.size == self.n and poly.flags["C_CONTIGUOUS"]
        getattr(self.host.lib, variant)(C.byref(self.c), poly.ctypes.data_as(u32p), degree)
        return poly

    def additive_fft_batch(self, polys: np.ndarray, num_terms: int) -> np.ndarray:
        polys, pp = _u32(polys)
        B = polys.shape[0]
        assert polys.shape[1] == self.n
        out = np.zeros((B, self.n), np.uint32)
        self.host._check(self.host.lib.additive_fft_batch(
            C.byref(self.c), pp, out.ctypes.data_as(u32p), num_terms, B), "additive_fft_batch")
        return out[:, : self.n - 1]

    def multiplicative_FFT(self, data: np.ndarray) -> int:
        """in place on data[0..n-2]; returns the reference's status (1 = unknown factorisation)"""
        assert data.dtype == np.uint32 and data.size >= self.n - 1
        buf = np.zeros(self.n, np.uint32)
        return self.host.lib.multiplicative_FFT(C.byref(self.c), data.ctypes.data_as(u32p),
                                                buf.ctypes.data_as(u32p))

    def evaluate_polynomial_multiplicative_FFT(self, poly: np.ndarray, num_coeffs: int):
        poly, pp = _u32(poly)
        out = np.zeros(self.n, np.uint32)
        buf = np.zeros(self.n, np.uint32)
        rc = self.host.lib.evaluate_polynomial_multiplicative_FFT(
            C.byref(self.c), pp, out.ctypes.data_as(u32p), num_coeffs, buf.ctypes.data_as(u32p))
        return rc, out[: self.n - 1]

    def multiplicative_fft_batch(self, polys: np.ndarray, num_coeffs: int) -> np.ndarray:
        polys, pp = _u32(polys)
        B = polys.shape[0]
        out = np.zeros((B, self.n), np.uint32)
        self.host._check(self.host.lib.multiplicative_fft_batch(
            C.byref(self.c), pp, out.ctypes.data_as(u32p), num_coeffs, B), "multiplicative_fft_batch")
        return out[:, : self.n - 1]


class BchCode:
    """bch_gen_code (libbch/bch.c:30-54) and the decode entry points."""

    def __init__(self, host: Host, codeword_size: int, minimum_distance: int):
        self.host = host
        self.c = host.lib.bch_gen_code(codeword_size, minimum_distance)
        if not self.c.m_packed_gen_poly:
            raise ValueError("code could not be created (bch_gen_code returned no generator)")
        self.length = self.c.m_length
        self.distance = self.c.m_distance
        self.t = (self.distance - 1) // 2
        self.m = self.c.m_field.m
        self.n = self.c.m_field.n
        self.gen_degree = self.c.m_gen_poly_degree
        self.words = (self.length - 1) // 64 + 1
        self.gen = np.ctypeslib.as_array(self.c.m_packed_gen_poly,
                                         (self.gen_degree // 64 + 1,)).astype(np.uint64)
        self.exp = np.ctypeslib.as_array(self.c.m_field.exp, (self.n,))
        self.log = np.ctypeslib.as_array(self.c.m_field.log, (self.n,))
        self._buf = None

    @property
    def data_bits(self):
        return self.length - self.gen_degree

    def encode(self, bits) -> np.ndarray:
        bits, bp = _u32(bits)
        cw = np.zeros(self.words, np.uint64)
        if self.host.lib.bch_encode(C.byref(self.c), cw.ctypes.data_as(ulp), bp, bits.size):
            raise ValueError("too many data bits")
        return cw

    def encode_batch(self, packed_data, data_bits: int) -> np.ndarray:
        """bch_encode_batch: packed_data [B, data_words] uint64 (bit k of an item = bit k % 64 of
        word k // 64) -> codewords [B, W]"""
        d = np.ascontiguousarray(packed_data, np.uint64)
        d = d.reshape(-1, d.shape[-1])
        out = np.zeros((d.shape[0], self.words), np.uint64)
        self.host._check(self.host.lib.bch_encode_batch(
            C.byref(self.c), out.ctypes.data_as(ulp), d.ctypes.data_as(ulp), d.shape[1], data_bits,
            d.shape[0]), "bch_encode_batch")
        return out

    def buffers(self):
        if self._buf is None:
            self._buf = self.host.lib.bch_alloc_buffers(C.byref(self.c))
        return self._buf

    def eval_syndrome(self, word):
        w = np.ascontiguousarray(word, np.uint64)
        b = self.buffers()
        flag = self.host.lib.bch_eval_syndrome(C.byref(self.c), C.byref(b), w.ctypes.data_as(ulp))
        return flag, np.ctypeslib.as_array(b.m_syndrome, (self.distance,)).copy()

    def compute_elp(self, syndrome):
        b = self.buffers()
        s = np.ascontiguousarray(syndrome, np.uint32)
        C.memmove(b.m_syndrome, s.ctypes.data, s.size * 4)
        L = self.host.lib.bch_compute_elp(C.byref(self.c), C.byref(b))
        return L, np.ctypeslib.as_array(b.m_C_elp, (self.distance + 1,))[: L + 1].copy()

    def locate_errors(self, elp, L):
        b = self.buffers()
        e = np.zeros(self.distance + 1, np.uint32)
        e[: len(elp)] = elp
        C.memmove(b.m_C_elp, e.ctypes.data, e.size * 4)
        cnt = self.host.lib.bch_locate_errors(C.byref(self.c), C.byref(b), L)
        return np.ctypeslib.as_array(b.m_error, (self.distance + 1,))[:cnt].copy()

    def decode(self, word):
        """bch_decode: returns (ok, corrected, word) -- `word` is decoded in a copy"""
        w = np.ascontiguousarray(word, np.uint64).copy()
        corr = C.c_uint32(0)
        ok = self.host.lib.bch_decode(C.byref(self.c), w.ctypes.data_as(ulp), None, C.byref(corr))
        return ok, corr.value, w

    def decode_batch(self, words):
        """bch_decode_batch: returns (ok[B], corrected[B], words[B, W]) on a copy of `words`"""
        w = np.ascontiguousarray(words, np.uint64).reshape(-1, self.words).copy()
        B = w.shape[0]
        ok = np.zeros(B, np.uint8)
        corr = np.zeros(B, np.uint32)
        self.host._check(self.host.lib.bch_decode_batch(
            C.byref(self.c), w.ctypes.data_as(ulp), B, ok.ctypes.data_as(C.POINTER(C.c_ubyte)),
            corr.ctypes.data_as(u32p)), "bch_decode_batch")
        return ok, corr, w


# ------------------------------------------------------------------------------------------
# device-resident layer (include/bchfft_cuda.h)
# ------------------------------------------------------------------------------------------
class DeviceField:
    """bchfft_field_create on `device`, from a FiniteField's tables."""

    def __init__(self, field: FiniteField, device: int = 0, lib=None):
        self.l = lib or _lib.lib()
        self.field = field
        self.device = device
        self.h = C.c_void_p()
        exp = np.ascontiguousarray(field.exp, np.uint32)
        log = np.ascontiguousarray(field.log, np.uint32)
        _lib.check(self.l, self.l.bchfft_field_create(device, field.m, exp.ctypes.data,
                                                      log.ctypes.data, C.byref(self.h)),
                   "bchfft_field_create")

    def additive_fft_dev(self, d_polys, d_results, num_terms, batch, stream=None,
                         d_natural=None, stride=None):
        n = self.field.n
        stride = stride or n
        _lib.check(self.l, self.l.bchfft_additive_fft_dev(
            self.h, d_polys, stride, d_results, stride, d_natural, stride, num_terms, batch,
            stream), "bchfft_additive_fft_dev")

    def additive_fft_host(self, polys, num_terms, want_natural=False):
        polys = np.ascontiguousarray(polys, np.uint32)
        B, n = polys.shape
        res = np.zeros((B, n), np.uint32)
        nat = np.zeros((B, n), np.uint32) if want_natural else None
        _lib.check(self.l, self.l.bchfft_additive_fft_host(
            self.h, polys.ctypes.data, n, res.ctypes.data, n,
            nat.ctypes.data if want_natural else None, n, num_terms, B),
            "bchfft_additive_fft_host")
        return (res[:, : n - 1], nat) if want_natural else res[:, : n - 1]

    def multiplicative_fft_dev(self, d_polys, d_evals, num_coeffs, batch, stream=None):
        n = self.field.n
        _lib.check(self.l, self.l.bchfft_multiplicative_fft_dev(
            self.h, d_polys, n, d_evals, n, num_coeffs, batch, stream),
            "bchfft_multiplicative_fft_dev")

    def sync(self):
        _lib.check(self.l, self.l.bchfft_field_sync(self.h), "bchfft_field_sync")

    def close(self):
        if self.h:
            self.l.bchfft_field_destroy(self.h)
            self.h = C.c_void_p()


class DeviceCode:
    """bchfft_code_create for a BchCode on a DeviceField."""

    def __init__(self, dfield: DeviceField, code: BchCode):
        self.l = dfield.l
        self.dfield = dfield
        self.code = code
        self.h = C.c_void_p()
        gen = np.ascontiguousarray(code.gen, np.uint64)
        _lib.check(self.l, self.l.bchfft_code_create(dfield.h, code.distance, code.length,
                                                     gen.ctypes.data, code.gen_degree,
                                                     C.byref(self.h)), "bchfft_code_create")

    def decode_dev(self, d_words, batch, d_ok, d_corrected, stream=None):
        _lib.check(self.l, self.l.bchfft_bch_decode_dev(self.h, d_words, batch, d_ok, d_corrected,
                                                        stream), "bchfft_bch_decode_dev")

    def encode_dev(self, d_data, data_words, data_bits, d_codewords, batch, stream=None):
        _lib.check(self.l, self.l.bchfft_bch_encode_dev(self.h, d_data, data_words, data_bits, d_codewords,
                                                        batch, stream), "bchfft_bch_encode_dev")

    def decode_host(self, words):
        w = np.ascontiguousarray(words, np.uint64).reshape(-1, self.code.words).copy()
        B = w.shape[0]
        ok = np.zeros(B, np.uint8)
        corr = np.zeros(B, np.uint32)
        _lib.check(self.l, self.l.bchfft_bch_decode_host(self.h, w.ctypes.data, B, ok.ctypes.data,
                                                         corr.ctypes.data), "bchfft_bch_decode_host")
        return ok, corr, w

    def decode_host_inplace(self, words_ptr, batch, ok_ptr, corr_ptr):
        _lib.check(self.l, self.l.bchfft_bch_decode_host(self.h, words_ptr, batch, ok_ptr, corr_ptr),
                   "bchfft_bch_decode_host")

    def close(self):
        if self.h:
            self.l.bchfft_code_destroy(self.h)
            self.h = C.c_void_p()
