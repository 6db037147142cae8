"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the two CPU parity checkers.

* ``Port``  -- oracle/libbchfft_oracle.so, the C restatement in oracle/bchfft_oracle.c.
* ``Ref``   -- oracle/_ref/libref_oracle.so, the UNMODIFIED reference sources compiled from
              /root/reference by oracle/Makefile (present in the build container and, as a
              prebuilt file that travels with the repo, on the GPU box).

Both expose the same numpy-level interface so tests can run one against the other and
either against the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libbchfft_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def build(quiet: bool = True) -> None:
    """Compile the checkers (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _p32(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def fnv64(arr: np.ndarray) -> int:
    h = 1469598103934665603
    for b in np.ascontiguousarray(arr).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


# --------------------------------------------------------------------------------------
# the restatement
# --------------------------------------------------------------------------------------
class _OrcField(C.Structure):
    _fields_ = [("m", C.c_uint32), ("n", C.c_uint32), ("order", C.c_uint32),
                ("exp", u32p), ("log", u32p)]


class _OrcCode(C.Structure):
    _fields_ = [("f", _OrcField), ("distance", C.c_uint32), ("gen_degree", C.c_uint32),
                ("length", C.c_uint32), ("gen", u64p)]


class Port:
    name = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = self.lib = C.CDLL(PORT_SO)
        L.orc_field_init.argtypes = [C.POINTER(_OrcField), C.c_uint32]
        L.orc_mul.argtypes = [C.POINTER(_OrcField), C.c_uint32, C.c_uint32]
        L.orc_mul.restype = C.c_uint32
        L.orc_bch_words.restype = C.c_size_t
        L.orc_bch_elp.restype = C.c_uint32
        L.orc_bch_locate.restype = C.c_uint32
        L.orc_bch_decode.restype = C.c_uint32
        L.orc_bch_syndrome.restype = C.c_int32
        L.orc_fnv64.restype = C.c_uint64
        L.orc_bch_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.c_void_p]
        self._fields = {}

    def field(self, m):
        if m not in self._fields:
            f = _OrcField()
            if self.lib.orc_field_init(C.byref(f), m):
                raise ValueError(f"unsupported field degree {m}")
            self._fields[m] = f
        return self._fields[m]

    def tables(self, m):
        f = self.field(m)
        n = 1 << m
        return (np.ctypeslib.as_array(f.exp, (n,)).copy(),
                np.ctypeslib.as_array(f.log, (n,)).copy())

    def mul(self, m, a, b):
        return self.lib.orc_mul(C.byref(self.field(m)), a, b)

    def subspace_polys(self, m):
        coef = np.zeros(m * (m + 1), np.uint32)
        y = np.zeros(m, np.uint32)
        self.lib.orc_subspace_polys(C.byref(self.field(m)), _p32(coef), _p32(y))
        return coef.reshape(m, m + 1), y

    def additive_fft(self, m, poly, num_terms):
        """returns (result[n-1] in exp order, clobbered poly[n] = natural-order values)"""
        n = 1 << m
        p = np.ascontiguousarray(poly, np.uint32).copy()
        assert p.size == n
        out = np.zeros(n - 1, np.uint32)
        self.lib.orc_eval_additive_fft(C.byref(self.field(m)), _p32(p), _p32(out),
                                       C.c_uint32(num_terms))
        return out, p

    def mult_fft(self, m, data):
        n = 1 << m
        d = np.ascontiguousarray(data, np.uint32).copy()
        assert d.size >= n - 1
        s = np.zeros(n, np.uint32)
        rc = self.lib.orc_multiplicative_fft(C.byref(self.field(m)), _p32(d), _p32(s))
        if rc:
            raise ValueError("unknown factorisation")
        return d[: n - 1]

    def mult_eval(self, m, poly, num_coeffs):
        n = 1 << m
        p = np.ascontiguousarray(poly, np.uint32)
        out = np.zeros(n, np.uint32)
        s = np.zeros(n, np.uint32)
        rc = self.lib.orc_eval_multiplicative_fft(C.byref(self.field(m)), _p32(p), _p32(out),
                                                  C.c_uint32(num_coeffs), _p32(s))
        if rc:
            raise ValueError("unknown factorisation")
        return out[: n - 1]

    def code(self, length, t, d=None):
        return PortCode(self, length, t, d)


class PortCode:
    def __init__(self, port, length, t, d=None):
        """d: designed distance when it is not 2t+1 (even distances: t = (d-1)//2)"""
        self.port, self.lib = port, port.lib
        self.c = _OrcCode()
        d = 2 * t + 1 if d is None else d
        t = (d - 1) // 2
        if self.lib.orc_bch_gen(C.byref(self.c), length, d):
            raise ValueError("code could not be created")
        self.length, self.t, self.d = length, t, d
        self.m = self.c.f.m
        self.gen_degree = self.c.gen_degree
        self.words = (length - 1) // 64 + 1
        self.gen = np.ctypeslib.as_array(self.c.gen, (self.gen_degree // 64 + 1,)).copy()

    def encode(self, bits):
        b = np.ascontiguousarray(bits, np.uint32)
        cw = np.zeros(self.words, np.uint64)
        if self.lib.orc_bch_encode(C.byref(self.c), _p64(cw), _p32(b), C.c_uint32(b.size)):
            raise ValueError("too many data bits")
        return cw

    def syndrome(self, word):
        w = np.ascontiguousarray(word, np.uint64)
        syn = np.zeros(self.d + 1, np.uint32)
        used = C.c_int(0)
        flag = self.lib.orc_bch_syndrome(C.byref(self.c), _p64(w), _p32(syn), C.byref(used))
        return flag, syn[: self.d], bool(used.value)

    def elp(self, syn):
        s = np.ascontiguousarray(syn, np.uint32)
        bufs = np.zeros(3 * (self.d + 1), np.uint32)
        elp = u32p()
        L = self.lib.orc_bch_elp(C.byref(self.c), _p32(s), _p32(bufs), C.byref(elp))
        off = (C.addressof(elp.contents) - bufs.ctypes.data) // 4
        return L, bufs[off: off + L + 1].copy()

    def locate(self, elp, L):
        e = np.zeros(self.d + 1, np.uint32)
        e[: len(elp)] = elp
        pos = np.zeros(1 << self.m, np.uint32)
        cnt = self.lib.orc_bch_locate(C.byref(self.c), _p32(e), C.c_uint32(L), _p32(pos))
        return pos[:cnt].copy()

    def decode(self, word):
        w = np.ascontiguousarray(word, np.uint64).copy()
        corr = C.c_uint32(0)
        ok = self.lib.orc_bch_decode(C.byref(self.c), _p64(w), C.byref(corr))
        return ok, corr.value, w

    def decode_batch(self, words):
        w = np.ascontiguousarray(words, np.uint64).copy().reshape(-1, self.words)
        ok = np.zeros(w.shape[0], np.uint8)
        corr = np.zeros(w.shape[0], np.uint32)
        self.lib.orc_bch_decode_batch(C.byref(self.c), w.ctypes.data, w.shape[0],
                                      ok.ctypes.data, corr.ctypes.data)
        return ok, corr, w


# --------------------------------------------------------------------------------------
# the unmodified reference (struct layouts: include/finite_field.h:13-26, include/bch.h:10-36)
# --------------------------------------------------------------------------------------
class _RefField(C.Structure):
    _fields_ = [("m", C.c_uint32), ("n", C.c_uint32), ("logzero", C.c_uint32),
                ("exp", u32p), ("log", u32p)]


class _RefCode(C.Structure):
    _fields_ = [("m_field", _RefField), ("m_distance", C.c_uint32),
                ("m_packed_gen_poly", u64p), ("m_gen_poly_degree", C.c_uint32),
                ("m_length", C.c_uint32)]


class _RefBuffers(C.Structure):
    _fields_ = [("m_syndrome", u32p), ("m_buffer_1", u32p), ("m_buffer_2", u32p),
                ("m_A", u32p), ("m_B", u32p), ("m_C_elp", u32p), ("m_error", u32p)]


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class Ref:
    name = "reference"

    def __init__(self):
        if not ref_available():
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.create_binary_finite_field.restype = _RefField
        L.create_binary_finite_field.argtypes = [C.c_uint32]
        L.multiply.restype = C.c_uint32
        L.multiply.argtypes = [C.POINTER(_RefField), C.c_uint32, C.c_uint32]
        L.bch_gen_code.restype = _RefCode
        L.bch_gen_code.argtypes = [C.c_uint32, C.c_uint32]
        L.bch_alloc_buffers.restype = _RefBuffers
        L.bch_decode.restype = C.c_uint32
        L.bch_compute_elp.restype = C.c_uint32
        L.bch_locate_errors.restype = C.c_uint32
        L.bch_eval_syndrome.restype = C.c_int32
        L.ref_bch_decode_batch_mt.restype = C.c_double
        L.ref_bch_decode_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                              C.c_void_p, C.c_int]
        L.ref_additive_fft_batch_mt.restype = C.c_double
        L.ref_additive_fft_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                C.c_size_t, C.c_int]
        L.ref_multiplicative_fft_batch_mt.restype = C.c_double
        L.ref_multiplicative_fft_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.ref_bch_encode_batch_mt.restype = C.c_double
        L.ref_bch_encode_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p,
                                              C.c_size_t, C.c_int]
        self._fields = {}

    def multiplicative_fft_batch_mt(self, m, polys, nthreads):
        """multi-threaded batch through the reference's multiplicative_FFT (in place on a copy);
        returns (evals[B, n-1], wall seconds)"""
        n = 1 << m
        p = np.ascontiguousarray(polys, np.uint32).copy().reshape(-1, n)
        dt = self.lib.ref_multiplicative_fft_batch_mt(C.byref(self.field(m)), p.ctypes.data, p.shape[0], nthreads)
        if dt < 0:
            raise ValueError("unknown factorisation")
        return p[:, : n - 1], dt

    def additive_fft_batch_mt(self, m, polys, num_terms, nthreads):
        """multi-threaded batch through the reference's evaluate_polynomial_additive_FFT;
        returns (results[B, n-1], wall seconds)"""
        n = 1 << m
        p = np.ascontiguousarray(polys, np.uint32).copy().reshape(-1, n)
        out = np.zeros_like(p)
        dt = self.lib.ref_additive_fft_batch_mt(C.byref(self.field(m)), p.ctypes.data,
                                                out.ctypes.data, num_terms, p.shape[0], nthreads)
        return out[:, : n - 1], dt

    def field(self, m):
        if m not in self._fields:
            f = self.lib.create_binary_finite_field(m)
            if not f.exp:
                raise ValueError(f"unsupported field degree {m}")
            self._fields[m] = f
        return self._fields[m]

    def tables(self, m):
        f = self.field(m)
        n = 1 << m
        return (np.ctypeslib.as_array(f.exp, (n,)).copy(),
                np.ctypeslib.as_array(f.log, (n,)).copy())

    def mul(self, m, a, b):
        return self.lib.multiply(C.byref(self.field(m)), a, b)

    def additive_fft(self, m, poly, num_terms, variant="evaluate_polynomial_additive_FFT"):
        n = 1 << m
        p = np.ascontiguousarray(poly, np.uint32).copy()
        assert p.size == n
        out = np.zeros(n, np.uint32)
        self.lib.evaluate_polynomial_additive_FFT(C.byref(self.field(m)), _p32(p), _p32(out),
                                                  C.c_uint32(num_terms))
        return out[: n - 1], p

    def additive_dft_variant(self, m, poly, degree, variant):
        """additive_DFT_opt / additive_DFT / additive_DFT_ref, in place, natural order"""
        p = np.ascontiguousarray(poly, np.uint32).copy()
        getattr(self.lib, variant)(C.byref(self.field(m)), _p32(p), C.c_uint32(degree))
        return p

    def mult_fft(self, m, data):
        n = 1 << m
        d = np.zeros(n, np.uint32)
        d[: n - 1] = np.asarray(data, np.uint32)[: n - 1]
        s = np.zeros(n, np.uint32)
        if self.lib.multiplicative_FFT(C.byref(self.field(m)), _p32(d), _p32(s)):
            raise ValueError("unknown factorisation")
        return d[: n - 1]

    def mult_eval(self, m, poly, num_coeffs):
        n = 1 << m
        p = np.ascontiguousarray(poly, np.uint32).copy()
        out = np.zeros(n, np.uint32)
        s = np.zeros(n, np.uint32)
        if self.lib.evaluate_polynomial_multiplicative_FFT(
                C.byref(self.field(m)), _p32(p), _p32(out), C.c_uint32(num_coeffs), _p32(s)):
            raise ValueError("unknown factorisation")
        return out[: n - 1]

    def code(self, length, t, d=None):
        return RefCode(self, length, t, d)


class RefCode:
    def __init__(self, ref, length, t, d=None):
        self.ref, self.lib = ref, ref.lib
        d = 2 * t + 1 if d is None else d
        t = (d - 1) // 2
        self.c = self.lib.bch_gen_code(length, d)
        if not self.c.m_packed_gen_poly:
            raise ValueError("code could not be created")
        self.length, self.t, self.d = length, t, d
        self.m = self.c.m_field.m
        self.gen_degree = self.c.m_gen_poly_degree
        self.words = (length - 1) // 64 + 1
        self.gen = np.ctypeslib.as_array(self.c.m_packed_gen_poly,
                                         (self.gen_degree // 64 + 1,)).copy()
        self.buf = self.lib.bch_alloc_buffers(C.byref(self.c))
        self._bm_bytes = (self.d + 1) * 4

    def _zero_bm(self):
        # the reference mallocs these (bch.c:568-575); zero them so >t-error behaviour is
        # deterministic
        for p in (self.buf.m_A, self.buf.m_B, self.buf.m_C_elp, self.buf.m_error):
            C.memset(p, 0, self._bm_bytes)

    def encode(self, bits):
        b = np.ascontiguousarray(bits, np.uint32)
        cw = np.zeros(self.words, np.uint64)
        if self.lib.bch_encode(C.byref(self.c), _p64(cw), _p32(b), C.c_uint32(b.size)):
            raise ValueError("too many data bits")
        return cw

    def syndrome(self, word):
        w = np.ascontiguousarray(word, np.uint64).copy()
        flag = self.lib.bch_eval_syndrome(C.byref(self.c), C.byref(self.buf), _p64(w))
        syn = np.ctypeslib.as_array(self.buf.m_syndrome, (self.d,)).copy()
        f = self.c.m_field
        used_fft = (self.d * self.gen_degree / f.n) > 0.5 * f.m * f.m
        return flag, syn, used_fft

    def elp(self, syn):
        s = np.ascontiguousarray(syn, np.uint32)
        C.memmove(self.buf.m_syndrome, s.ctypes.data, s.size * 4)
        self._zero_bm()
        L = self.lib.bch_compute_elp(C.byref(self.c), C.byref(self.buf))
        return L, np.ctypeslib.as_array(self.buf.m_C_elp, (L + 1,)).copy()

    def locate(self, elp, L):
        e = np.zeros(self.d + 1, np.uint32)
        e[: len(elp)] = elp
        C.memmove(self.buf.m_C_elp, e.ctypes.data, e.size * 4)
        cnt = self.lib.bch_locate_errors(C.byref(self.c), C.byref(self.buf), C.c_uint32(L))
        return np.ctypeslib.as_array(self.buf.m_error, (max(cnt, 1),))[:cnt].copy()

    def _decode_one(self, word):
        # Decode in a private buffer long enough for bit `order`: for shortened codes the
        # reference's bit-0 quirk (bch.c:485 check commented out) flips a bit beyond the packed
        # codeword; here that write lands in padding and is dropped.
        pad = np.zeros(max(self.words, (1 << self.m) // 64 + 1), np.uint64)
        pad[: self.words] = word
        self._zero_bm()
        corr = C.c_uint32(0)
        ok = self.lib.bch_decode(C.byref(self.c), _p64(pad), C.byref(self.buf), C.byref(corr))
        return ok, corr.value, pad[: self.words].copy()

    def decode(self, word):
        return self._decode_one(np.ascontiguousarray(word, np.uint64))

    def decode_batch_mt(self, words, nthreads):
        """multi-threaded batch through the reference's bch_decode (full-length codes only: no
        padding per word); returns (ok, corrected, words, wall seconds)"""
        assert self.words * 64 > (1 << self.m) - 1, "shortened code: use decode_batch"
        w = np.ascontiguousarray(words, np.uint64).copy().reshape(-1, self.words)
        ok = np.zeros(w.shape[0], np.uint8)
        corr = np.zeros(w.shape[0], np.uint32)
        dt = self.lib.ref_bch_decode_batch_mt(C.byref(self.c), w.ctypes.data, w.shape[0],
                                              ok.ctypes.data, corr.ctypes.data, nthreads)
        return ok, corr, w, dt

    def decode_batch(self, words):
        w = np.ascontiguousarray(words, np.uint64).copy().reshape(-1, self.words)
        ok = np.zeros(w.shape[0], np.uint8)
        corr = np.zeros(w.shape[0], np.uint32)
        for i in range(w.shape[0]):
            ok[i], corr[i], w[i] = self._decode_one(w[i])
        return ok, corr, w

    def encode_batch_mt(self, packed_data, data_bits, nthreads):
        """multi-threaded batch through the reference's bch_encode from packed data bits
        ([B, data_words] uint64); returns (codewords [B, W], wall seconds)"""
        d = np.ascontiguousarray(packed_data, np.uint64)
        out = np.zeros((d.shape[0], self.words), np.uint64)
        dt = self.lib.ref_bch_encode_batch_mt(C.byref(self.c), d.ctypes.data, d.shape[1], data_bits,
                                              out.ctypes.data, d.shape[0], nthreads)
        return out, dt
