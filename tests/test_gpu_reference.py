"""Parity of the CUDA path against the UNMODIFIED reference (oracle/_ref/libref_oracle.so: the
reference's own sources compiled by oracle/Makefile; it travels to the GPU box as a prebuilt file).
test_gpu_parity.py compares with the C restatement; here the chain is one hop shorter -- CUDA output
against the reference's output on the same inputs -- for every BASELINE configuration and for the
parameter ranges the reference nominally supports (fields up to GF(2^25), the (1048575, t = 100)
code of SURVEY appendix C).  Everything is bit-exact."""
import ctypes as C

import numpy as np
import pytest

from bch_fft_b200 import _lib, synth
from tests.common import flip, fnv64

pytestmark = pytest.mark.gpu


def _cores():
    import os
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def test_reference_library_is_loaded(gpu_libs, ref):
    assert ref.mul(13, 3, 7) == 9
    maps = open("/proc/self/maps").read()
    assert "libref_oracle.so" in maps and "libbchfft_cuda.so" in maps


# ---- additive FFT ----------------------------------------------------------------------------
@pytest.mark.parametrize("m,B", [(13, 70), (16, 40), (17, 6), (18, 5), (19, 4), (20, 3), (21, 2)])
def test_additive_fft_vs_reference(gpu_libs, ref, m, B):
    """evaluate_polynomial_additive_FFT (additive_fft.c:13-21): full degree, truncated degrees and the
    clobbered natural-order array, for every BASELINE field and the fields between and above them"""
    _, host = gpu_libs
    n = 1 << m
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, B, 0xA0 + m)
    for terms in (n - 1, 1000 % n, n // 2, n // 2 - 1, 0, 1):
        got = f.additive_fft_batch(polys, terms)
        want, _ = ref.additive_fft_batch_mt(m, polys, terms, _cores())
        assert (got == want).all(), (m, terms)
    # single-call entry point: po_result and the clobbered p_poly
    p = polys[0].copy()
    res = f.evaluate_polynomial_additive_FFT(p, n - 1)
    rres, rnat = ref.additive_fft(m, polys[0], n - 1)
    assert (res == rres).all() and (p == rnat).all()


def _eval_points(exp, log, poly, idx):
    """P(x^i) for i in idx by direct summation in the log domain (numpy, independent of any FFT)"""
    order = len(exp) - 1
    k = np.nonzero(poly)[0].astype(np.uint64)
    lc = log[poly[k]].astype(np.uint64)
    out = []
    for i in idx:
        e = (lc + (k * np.uint64(i)) % np.uint64(order)) % np.uint64(order)
        out.append(int(np.bitwise_xor.reduce(exp[e.astype(np.int64)])))
    return np.array(out, np.uint32)


@pytest.mark.parametrize("m", [22, 24, 25])
def test_additive_fft_large_fields(gpu_libs, ref, m):
    """GF(2^22) .. GF(2^25), the top of the reference's field table (finite_field.c:9-18): truncated
    evaluation against the reference (its low-degree shortcut keeps the CPU time in seconds), full
    degree against direct evaluation at sampled points with the reference's own exp / log tables"""
    _, host = gpu_libs
    n = 1 << m
    f = host.create_binary_finite_field(m)
    rexp, rlog = ref.tables(m)
    assert (f.exp == rexp).all() and (f.log == rlog).all()
    polys = synth.random_polys(m, 2, 0xB0 + m)
    got = f.additive_fft_batch(polys, 300)
    want, _ = ref.additive_fft_batch_mt(m, polys, 300, 2)
    assert (got == want).all()
    got = f.additive_fft_batch(polys, n - 1)
    idx = [0, 1, 2, n - 2, n // 3, 12345678 % (n - 1), (1 << (m - 1)) + 5]
    for b in range(2):
        assert (got[b, idx] == _eval_points(rexp, rlog, polys[b], idx)).all(), (m, b)
    # linearity: FFT(a ^ b) = FFT(a) ^ FFT(b)
    mix = (polys[0] ^ polys[1])[None, :].copy()
    assert (f.additive_fft_batch(mix, n - 1)[0] == (got[0] ^ got[1])).all()


def test_additive_fft_gf2_20_full_batch_vs_reference(gpu_libs, ref):
    """BASELINE configs[3]: 256 polynomials over GF(2^20), 1 GiB of coefficients; 32 of
    them (every 8th, plus neighbours of the slab boundaries) against the reference on all host cores,
    the demo input's Appendix C fingerprint, and additive == multiplicative on the device"""
    cabi, host = gpu_libs
    m, n, B = 20, 1 << 20, 256
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, B, 0xC4)
    polys[255] = synth.demo_poly(m)
    got = f.additive_fft_batch(polys, n - 1)
    assert fnv64(got[255]) == 0x43310189501ffec7
    pick = sorted(set(range(0, B, 8)) | {31, 32, 33, 223, 224, 254})[:40]
    want, _ = ref.additive_fft_batch_mt(m, polys[pick], n - 1, _cores())
    assert (got[pick] == want).all()
    mult = f.multiplicative_fft_batch(polys[:33], n - 1)
    assert (mult == got[:33]).all()


def test_additive_fft_gf2_16_full_batch_vs_reference(gpu_libs, ref):
    """BASELINE configs[0] batched: 4096 polynomials over GF(2^16); 128 of them against
    the reference, the demo input's fingerprint"""
    _, host = gpu_libs
    m, n, B = 16, 1 << 16, 4096
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, B, 0xC1)
    polys[4095] = synth.demo_poly(m)
    got = f.additive_fft_batch(polys, n - 1)
    assert fnv64(got[4095]) == 0xda6143e12631d2df
    pick = sorted(set(range(0, B, 33)) | {31, 32, 4064, 4094})
    want, _ = ref.additive_fft_batch_mt(m, polys[pick], n - 1, _cores())
    assert (got[pick] == want).all()


# ---- multiplicative FFT ------------------------------------------------------------------------
@pytest.mark.parametrize("m", [8, 9, 10, 11, 12, 14, 15, 16, 18, 20])
def test_multiplicative_fft_vs_reference(gpu_libs, ref, m):
    """multiplicative_FFT / evaluate_polynomial_multiplicative_FFT (multiplicative_fft.c:106-324) for
    every factorisation of the reference's table with small prime factors, full and truncated"""
    _, host = gpu_libs
    n = 1 << m
    f = host.create_binary_finite_field(m)
    B = 36 if m <= 16 else 3
    polys = synth.random_polys(m, B, 0xC5 + m)
    full = f.multiplicative_fft_batch(polys, n - 1)
    for b in (0, B - 1):
        assert (full[b] == ref.mult_fft(m, polys[b, : n - 1])).all(), (m, b)
    for nc in (1, 2, 15, 255, (n - 1) // 3, (n - 1) // 5 + 1, n - 2):
        nc = min(nc, n - 1)
        got = f.multiplicative_fft_batch(polys, nc)
        for b in (0, B - 1):
            assert (got[b] == ref.mult_eval(m, polys[b], nc)).all(), (m, nc, b)


def test_multiplicative_fft_gf2_16_full_batch(gpu_libs, ref):
    """BASELINE configs[4]: 4096 polynomials over GF(2^16) -- equal to the additive
    path's po_result for the whole batch (fft/test.c:147 relation), 24 of them against the reference"""
    _, host = gpu_libs
    m, n, B = 16, 1 << 16, 4096
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, B, 0xC5)
    mult = f.multiplicative_fft_batch(polys, n - 1)
    add = f.additive_fft_batch(polys, n - 1)
    assert (mult == add).all()
    for b in list(range(0, B, 190)) + [31, 32, 4095]:
        assert (mult[b] == ref.mult_fft(m, polys[b, : n - 1])).all(), b


# ---- decode -------------------------------------------------------------------------------------
def _words_for(rc, n, t, B, rng):
    pool = [rc.encode(rng.integers(0, 2, n - rc.gen_degree, dtype=np.uint32)) for _ in range(4)]
    words = np.zeros((B, rc.words), np.uint64)
    for b in range(B):
        w = pool[b % 4].copy()
        ne = [0, 1, t, t // 2, t + 1, t + 2, t + 3][b % 7] if b < 28 else int(rng.integers(0, t + 1))
        pos = rng.choice(n, min(ne, n), replace=False)
        if b % 5 == 0 and ne:
            pos[0] = 0                                # bit-0 quirk (bch.c:479-488)
        for p in set(int(x) for x in pos):
            flip(w, p)
        words[b] = w
    return words


@pytest.mark.parametrize("n,t,B", [(8191, 16, 64), (8191, 64, 48), (8191, 200, 40), (65535, 100, 40),
                                   (65535, 1000, 30), (1048575, 100, 34), (131071, 20, 34),
                                   (262143, 10, 33)])
def test_decode_stages_vs_reference(gpu_libs, ref, n, t, B):
    """bch_eval_syndrome -> bch_compute_elp -> bch_locate_errors -> bch_decode (bch.c:228-562) on every
    BASELINE code and on fields above GF(2^16) -- (1048575, 100) is the appendix C code over GF(2^20),
    which runs k_syndrome<20>, k_bm<20>, k_locate<20> -- word by word against the reference"""
    _, host = gpu_libs
    code = host.bch_gen_code(n, 2 * t + 1)
    rc = ref.code(n, t)
    assert code.gen_degree == rc.gen_degree and (code.gen == rc.gen).all()
    if (n, t) == (1048575, 100):
        assert f"{fnv64(code.gen):016x}" == "38dc62bb051634f8"
    words = _words_for(rc, n, t, B, np.random.default_rng(n * 3 + t))
    ok, corr, out = code.decode_batch(words)
    rok, rcorr, rout = rc.decode_batch(words)
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()
    for b in list(range(0, 14)) + [B - 1]:
        fl, syn = code.eval_syndrome(words[b])
        rfl, rsyn, _ = rc.syndrome(words[b])
        assert fl == rfl and (syn == rsyn).all(), b
        L, elp = code.compute_elp(syn)
        rL, relp = rc.elp(syn)
        assert L == rL and (elp == relp).all(), b
        if L <= t:
            pos = code.locate_errors(elp, L)
            assert (pos == rc.locate(relp, rL)).all(), b
        o, c, w = code.decode(words[b])
        assert (o, c) == (int(rok[b]), int(rcorr[b])) and (w == rout[b]).all(), b


@pytest.mark.parametrize("n,d", [(8191, 34), (65535, 12)])
def test_even_designed_distance_vs_reference(gpu_libs, ref, n, d):
    """any designed distance the reference accepts (bch.c:37): even d = 2t+2"""
    _, host = gpu_libs
    t = (d - 1) // 2
    code = host.bch_gen_code(n, d)
    rc = ref.code(n, t, d=d)
    assert (code.gen == rc.gen).all()
    words = _words_for(rc, n, t, 60, np.random.default_rng(n + d))
    ok, corr, out = code.decode_batch(words)
    rok, rcorr, rout = rc.decode_batch(words)
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()


@pytest.mark.parametrize("t,B", [(16, 40000), (64, 20000), (200, 12000)])
def test_batched_decode_vs_reference_at_scale(gpu_libs, ref, t, B):
    """BASELINE configs[2], all three codes of the survey: (8191, t) for t = 16, 64, 200
    -- t >= 32 runs the general root-search kernel, t = 200 the reference's FFT paths -- with the
    bench's error mix (0..t errors, 1 % beyond capacity, bit 0 included), every word compared with the
    reference run on all host cores"""
    _, host = gpu_libs
    n = 8191
    code = host.bch_gen_code(n, 2 * t + 1)
    rc = ref.code(n, t)
    bits = (synth.splitmix64(0xC3, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([rc.encode(b) for b in bits.reshape(64, -1)])
    words, nerr = synth.corrupt_codewords(pool, n, t, B, 0xC3 + t, beyond_fraction=0.01)
    ok, corr, out = code.decode_batch(words)
    rok, rcorr, rout, _ = rc.decode_batch_mt(words, _cores())
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()
    assert 0.97 < ok.mean() < 1.0 and corr.max() >= t


@pytest.mark.parametrize("n,t", [(8191, 200), (65535, 1000), (8191, 16)])
def test_encode_batch_vs_reference(gpu_libs, ref, n, t):
    """bch_encode (bch.c:492-506) for generator degrees on both sides of 2048: (8191, 200) has
    deg g = 2444 and (65535, 1000) deg g = 15360 (warp-per-codeword kernel), (8191, 16) 208"""
    _, host = gpu_libs
    code = host.bch_gen_code(n, 2 * t + 1)
    rc = ref.code(n, t)
    k = code.data_bits
    rng = np.random.default_rng(n + t)
    for nbits in (k, k - 1, 65, 1):
        B = 70
        bits = rng.integers(0, 2, (B, nbits), dtype=np.uint32)
        bits[0] = 0
        bits[1] = 1
        packed = np.zeros((B, (nbits + 63) // 64), np.uint64)
        for w in range(packed.shape[1]):
            chunk = bits[:, 64 * w: 64 * w + 64].astype(np.uint64)
            packed[:, w] = (chunk << np.arange(chunk.shape[1], dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
        out = code.encode_batch(packed, nbits)
        for b in (0, 1, 2, 33, B - 1):
            assert (out[b] == rc.encode(bits[b])).all(), (n, t, nbits, b)


# ---- end-to-end result modes, device sets ----------------------------------------------------------
def _pinned_array(cabi, shape, dtype):
    import ctypes as C
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_pinned_alloc(nbytes, C.byref(p)), "pinned_alloc")
    buf = (C.c_ubyte * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def test_decode_patch_mode_on_pinned_host_memory(gpu_libs, ref):
    """bch_decode_batch on a page-locked array in its three result modes (BCHFFT_D2H_MODE): positions
    (verdicts + error positions come back, host threads flip the bits: the default for large batches),
    words (every word is copied back), patch (k_correct writes the changed words into the caller's array).
    All modes, and pageable memory, against the reference -- 50000 words incl. failures (left untouched),
    bit-0 quirk, two errors in one 64-bit word."""
    import ctypes as C
    import os
    cabi, host = gpu_libs
    n, t, B = 8191, 16, 50000
    code, rc = host.bch_gen_code(n, 2 * t + 1), ref.code(n, t)
    bits = (synth.splitmix64(0xD2, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([rc.encode(b) for b in bits.reshape(64, -1)])
    words, _ = synth.corrupt_codewords(pool, n, t, B, 0xD2, beyond_fraction=0.03)
    words[7] = pool[3]
    flip(words[7], 128), flip(words[7], 129), flip(words[7], 0)
    rok, rcorr, rout, _ = rc.decode_batch_mt(words, _cores())
    pw, hp = _pinned_array(cabi, words.shape, np.uint64)
    ulp = C.POINTER(C.c_ulong)
    try:
        for mode in ("positions", "patch", "words"):
            os.environ["BCHFFT_D2H_MODE"] = mode
            pw[:] = words
            ok = np.zeros(B, np.uint8)
            corr = np.zeros(B, np.uint32)
            rcode = host.lib.bch_decode_batch(C.byref(code.c), C.cast(hp.value, ulp), B,
                                              ok.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                              corr.ctypes.data_as(C.POINTER(C.c_uint32)))
            assert rcode == 0
            assert (ok == rok).all() and (corr == rcorr).all() and (pw == rout).all(), mode
    finally:
        os.environ.pop("BCHFFT_D2H_MODE", None)
        cabi.bchfft_pinned_free(hp)
    ok, corr, out = code.decode_batch(words)          # pageable numpy memory
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()


def test_in_library_multi_gpu_sharding(gpu_libs, ref):
    """SURVEY 8e inside the library: with several devices a batch call drives one host thread per GPU on
    contiguous ranges of the caller's single array (the array is the gather).  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    cabi, _ = gpu_libs
    ndev = cabi.bchfft_device_count()
    if ndev < 2:
        pytest.skip("one GPU visible: the multi-device driver is covered by the emulator test "
                    "test_emu_in_library_multi_device_sharding")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from bch_fft_b200 import api, synth\n"
        "from oracle import pyoracle\n"
        "host, R = api.Host(), pyoracle.Ref()\n"
        f"assert host.lib.bchfft_host_num_devices() == {ndev}\n"
        "n, t, B = 8191, 16, 100000\n"
        "code, rc = host.bch_gen_code(n, 2 * t + 1), R.code(n, t)\n"
        "bits = (synth.splitmix64(0xD3, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)\n"
        "pool = np.stack([rc.encode(b) for b in bits.reshape(64, -1)])\n"
        "words, _ = synth.corrupt_codewords(pool, n, t, B, 0xD3, beyond_fraction=0.02)\n"
        "ok, corr, out = code.decode_batch(words)\n"
        "rok, rcorr, rout, _ = rc.decode_batch_mt(words, 16)\n"
        "assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()\n"
        "f = host.create_binary_finite_field(16)\n"
        "polys = synth.random_polys(16, 512, 0xD4)\n"
        "res = f.additive_fft_batch(polys, 65535)\n"
        "pick = list(range(0, 512, 37)) + [511]\n"
        "want, _ = R.additive_fft_batch_mt(16, polys[pick], 65535, 16)\n"
        "assert (res[pick] == want).all()\n"
        "mul = f.multiplicative_fft_batch(polys, 65535)\n"
        "assert (mul == res).all()\n"
        "print('contexts', host.lib.bchfft_host_num_contexts())\n")
    env = dict(os.environ, BCHFFT_DEVICES="all", BCHFFT_MIN_PER_DEVICE="64")
    env.pop("BCHFFT_DEVICE", None)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert res.returncode == 0, res.stderr[-3000:]
    # per device: GF(2^13) field + code, GF(2^16) field
    assert f"contexts {3 * ndev}" in res.stdout, res.stdout
