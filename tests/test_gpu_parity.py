"""Parity tests proper: the CUDA path (through the C-ABI and through the drop-in host library)
against the oracle on seeded inputs, against the golden fixtures produced by the unmodified
reference, and -- at BASELINE.json's full sizes -- through size-independent properties.
Everything here is bit-exact (integer / GF(2^m) work)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bch_fft_b200 import _lib, api, synth
from tests.common import GOLDEN, additive_input, decode_input, flip, fnv64, low_degree_locators

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(cabi, port, m):
    exp, log = port.tables(m)
    h = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_field_create(0, m, exp.ctypes.data, log.ctypes.data, C.byref(h)), "field")
    return h


def _fft(cabi, h, polys, terms, natural=True):
    B, n = polys.shape
    res = np.full((B, n), 0xDEADBEEF, np.uint32)
    nat = np.zeros((B, n), np.uint32) if natural else None
    _lib.check(cabi, cabi.bchfft_additive_fft_host(h, polys.ctypes.data, n, res.ctypes.data, n,
                                                   nat.ctypes.data if natural else None, n, terms, B), "fft")
    assert (res[:, n - 1] == 0xDEADBEEF).all()
    return res[:, : n - 1], nat


def test_native_library_is_loaded(gpu_libs):
    cabi, host = gpu_libs
    assert cabi.bchfft_is_emulated() == 0 and cabi.bchfft_device_count() >= 1
    maps = open("/proc/self/maps").read()
    assert "libbchfft_cuda.so" in maps and "libbchfft_emu.so" not in maps


@pytest.mark.parametrize("m", [4, 8, 10, 13, 16])
def test_additive_fft_vs_oracle(gpu_libs, port, m):
    cabi, _ = gpu_libs
    n = 1 << m
    h = _field(cabi, port, m)
    rng = np.random.default_rng(100 + m)
    for B, terms in ((1, n - 1), (33, n - 1), (64, n - 1), (3, 0), (3, 1), (3, 2), (5, 5),
                     (34, n // 2 - 1), (7, n // 2), (2, n - 2), (2, n + 5), (5, 100 % n), (5, 1000 % n)):
        polys = rng.integers(0, n, (B, n), dtype=np.uint32)
        res, nat = _fft(cabi, h, polys, terms)
        check = range(B) if m <= 13 else sorted({0, B // 2, B - 1})
        for b in check:
            o, c = port.additive_fft(m, polys[b], min(terms, n - 1))
            assert (res[b] == o).all() and (nat[b] == c).all(), (m, B, terms, b)
    cabi.bchfft_field_destroy(h)


@pytest.mark.parametrize("rec", GOLDEN["additive"], ids=lambda r: f"m{r['m']}-terms{r['num_terms']}")
def test_additive_fft_golden(gpu_libs, port, rec):
    cabi, _ = gpu_libs
    h = _field(cabi, port, rec["m"])
    res, nat = _fft(cabi, h, additive_input(rec)[None, :].copy(), rec["num_terms"])
    assert f"{fnv64(res[0]):016x}" == rec["fnv_result"] and f"{fnv64(nat[0]):016x}" == rec["fnv_natural"]
    assert [int(x) for x in res[0, :8]] == rec["head"]
    cabi.bchfft_field_destroy(h)


def test_additive_fft_gf2_20_batch(gpu_libs, port):
    """BASELINE configs[3]: GF(2^20), transform larger than one SM's tile (multi-pass schedule)"""
    cabi, _ = gpu_libs
    m, n, B = 20, 1 << 20, 34
    h = _field(cabi, port, m)
    polys = synth.random_polys(m, B, 0xC4)
    res, _ = _fft(cabi, h, polys, n - 1, natural=False)
    for b in (0, 33):
        assert (res[b] == port.additive_fft(m, polys[b], n - 1)[0]).all()
    # linearity across the batch (size-independent property): FFT(a ^ b) = FFT(a) ^ FFT(b)
    mix = (polys[1] ^ polys[2])[None, :].copy()
    r2, _ = _fft(cabi, h, mix, n - 1, natural=False)
    assert (r2[0] == (res[1] ^ res[2])).all()
    cabi.bchfft_field_destroy(h)


def test_additive_fft_full_batch_properties(gpu_libs, port):
    """BASELINE configs[0] batched at full size (4096 polynomials over GF(2^16), 1 GiB):
    spot parity against the oracle + linearity + batch-position independence"""
    cabi, _ = gpu_libs
    m, n, B = 16, 1 << 16, 4096
    h = _field(cabi, port, m)
    polys = synth.random_polys(m, B, 0xC1)
    polys[7] = polys[5] ^ polys[6]
    polys[4095] = polys[3]
    res, _ = _fft(cabi, h, polys, n - 1, natural=False)
    for b in (0, 31, 32, 2047, 4095):
        assert (res[b] == port.additive_fft(m, polys[b], n - 1)[0]).all(), b
    assert (res[7] == (res[5] ^ res[6])).all()
    assert (res[4095] == res[3]).all()
    cabi.bchfft_field_destroy(h)


def _words_for(pc, n, t, B, rng):
    pool = [pc.encode(rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint32)) for _ in range(4)]
    words = np.zeros((B, pc.words), np.uint64)
    for b in range(B):
        w = pool[b % 4].copy()
        ne = [0, 1, t, t // 2, t + 1, t + 2, t + 3][b % 7] if b < 56 else int(rng.integers(0, t + 1))
        pos = rng.choice(n, min(ne, n), replace=False)
        if b % 5 == 0 and ne:
            pos[0] = 0
        for p in set(int(x) for x in pos):
            flip(w, p)
        words[b] = w
    return words


@pytest.mark.parametrize("n,t", [(15, 2), (63, 5), (255, 8), (255, 20), (1023, 10), (300, 3), (5000, 12),
                                 (8191, 8), (8191, 16), (8191, 64), (8191, 200), (65535, 16),
                                 (65535, 100), (65535, 1000),
                                 # fields with a periodic Cantor basis (m = 12: groups of 4 depths, m = 14: of 2)
                                 (4095, 12), (4095, 40), (16383, 10)])
def test_bch_stages_and_decode_vs_oracle(gpu_libs, port, n, t):
    cabi, _ = gpu_libs
    pc = port.code(n, t)
    h = _field(cabi, port, pc.m)
    c = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_code_create(h, 2 * t + 1, n, pc.gen.ctypes.data, pc.gen_degree,
                                             C.byref(c)), "code")
    B = 100 if t <= 200 else 40
    d = 2 * t + 1
    words = _words_for(pc, n, t, B, np.random.default_rng(n + t))
    syn = np.zeros((B, d), np.uint32)
    flags = np.zeros(B, np.int32)
    _lib.check(cabi, cabi.bchfft_bch_syndrome_host(c, words.ctypes.data, B, syn.ctypes.data,
                                                   flags.ctypes.data), "syndrome")
    elp = np.zeros((B, d + 1), np.uint32)
    L = np.zeros(B, np.uint32)
    _lib.check(cabi, cabi.bchfft_bch_elp_host(c, syn.ctypes.data, B, elp.ctypes.data, L.ctypes.data), "elp")
    posn = np.zeros((B, d + 1), np.uint32)
    cnt = np.zeros(B, np.uint32)
    _lib.check(cabi, cabi.bchfft_bch_locate_host(c, elp.ctypes.data, L.ctypes.data, B,
                                                 posn.ctypes.data, cnt.ctypes.data), "locate")
    for b in range(B):
        fl, s, used_fft = pc.syndrome(words[b])
        assert (syn[b] == s).all() and flags[b] == fl      # incl. the method-dependent entry 0
        Lr, er = pc.elp(syn[b])
        assert L[b] == Lr and (elp[b, : Lr + 1] == er).all()
        pr = pc.locate(elp[b, : Lr + 1], Lr)
        assert cnt[b] == len(pr) and (posn[b, : cnt[b]] == pr).all()
    ok = np.zeros(B, np.uint8)
    corr = np.zeros(B, np.uint32)
    out = words.copy()
    _lib.check(cabi, cabi.bchfft_bch_decode_host(c, out.ctypes.data, B, ok.ctypes.data,
                                                 corr.ctypes.data), "decode")
    rok, rcorr, rw = pc.decode_batch(words)
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rw).all()
    cabi.bchfft_code_destroy(c)
    cabi.bchfft_field_destroy(h)


@pytest.mark.parametrize("rec", GOLDEN["decode"], ids=lambda r: f"n{r['n']}-t{r['t']}")
def test_host_api_decode_golden(gpu_libs, rec):
    """reference-shaped host API (bch_gen_code / bch_encode / bch_decode[_batch]) against the
    fixtures generated by the unmodified reference"""
    _, host = gpu_libs
    code = host.bch_gen_code(rec["n"], 2 * rec["t"] + 1)
    grec = next(c for c in GOLDEN["codes"] if c["n"] == rec["n"] and c["t"] == rec["t"])
    assert code.gen_degree == grec["gen_degree"] and f"{fnv64(code.gen):016x}" == grec["fnv_gen"]
    pool, words = decode_input(rec, code.encode)
    assert f"{fnv64(pool):016x}" == rec["fnv_pool"] and f"{fnv64(words):016x}" == rec["fnv_in"]
    ok, corr, out = code.decode_batch(words)
    assert [int(x) for x in ok] == rec["ok"] and [int(x) for x in corr] == rec["corrected"]
    assert f"{fnv64(out):016x}" == rec["fnv_out"]
    for i in (0, 1, 2, len(words) - 1):
        o, c, w = code.decode(words[i])
        assert (o, c) == (rec["ok"][i], rec["corrected"][i]) and (w == out[i]).all()


def test_host_api_additive_single_call(gpu_libs, port):
    """evaluate_polynomial_additive_FFT / additive_DFT* drop-in semantics (additive_fft.c:13-21)"""
    _, host = gpu_libs
    f = host.create_binary_finite_field(16)
    poly = synth.demo_poly(16)
    p = poly.copy()
    got = f.evaluate_polynomial_additive_FFT(p, (1 << 16) - 1)
    assert fnv64(got) == 0xda6143e12631d2df and fnv64(p) == 0x4ce7d70654a732f3   # Appendix C
    rng = np.random.default_rng(4)
    poly = rng.integers(0, 1 << 16, 1 << 16, dtype=np.uint32)
    want_nat = port.additive_fft(16, poly, 700)[1]
    for variant in ("additive_DFT_opt", "additive_DFT", "additive_DFT_ref"):
        p = poly.copy()
        assert (f.additive_DFT(p, 700, variant) == want_nat).all()


def test_decode_full_batch_properties(gpu_libs, port):
    """BASELINE configs[2] at full size (10^6 words, n=8191, t=16): encode -> corrupt -> decode
    round trip.  Every word with <= t distinct errors (none at bit 0) must come back equal to its
    clean codeword with corrected = #errors; decoding twice is idempotent; a 20000-word sample is
    compared field by field with the oracle."""
    _, host = gpu_libs
    n, t, B = 8191, 16, 1_000_000
    code = host.bch_gen_code(n, 2 * t + 1)
    bits = (synth.splitmix64(0xC3, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([code.encode(b) for b in bits.reshape(64, -1)])
    # rebuild the same words without errors: same seed, t = 0 draws nothing
    r = synth.splitmix64(0xC3, B * 6).reshape(B, 6)
    clean = np.zeros((B, code.words), np.uint64)
    for k in range(4):
        clean ^= pool[(r[:, k] % np.uint64(64)).astype(np.int64)]
    words, nerr = synth.corrupt_codewords(pool, n, t, B, 0xC3, beyond_fraction=0.01)
    diff = words ^ clean
    ok, corr, out = code.decode_batch(words)
    sample = 20000
    pc = port.code(n, t)
    rok, rcorr, rout = pc.decode_batch(words[:sample])
    assert (ok[:sample] == rok).all() and (corr[:sample] == rcorr).all() and (out[:sample] == rout).all()
    # round trip on the full batch
    bit0 = (diff[:, 0] & np.uint64(1)) != 0
    tbl = np.array([bin(i).count("1") for i in range(256)], np.uint8)
    popc = tbl[diff.view(np.uint8).reshape(B, -1)].sum(axis=1, dtype=np.int64)
    good = (popc <= t) & ~bit0
    assert good.sum() > 0.9 * B
    assert (ok[good] == 1).all() and (corr[good] == popc[good]).all()
    assert (out[good] == clean[good]).all()
    assert (out[ok == 0] == words[ok == 0]).all()            # failures leave the word untouched
    ok2, corr2, out2 = code.decode_batch(out[good][:200000])
    assert (ok2 == 1).all() and (corr2 == 0).all() and (out2 == out[good][:200000]).all()


@pytest.mark.parametrize("n,t", [(15, 2), (255, 8), (1023, 10), (5000, 12), (8191, 16), (8191, 64), (65535, 16)])
def test_encode_batch_vs_reference_encoder(gpu_libs, port, n, t):
    """device encoder (bch_encode_batch) == bch_encode per item == the oracle's encoder, for full,
    short and ragged data lengths"""
    from tests.test_emu_kernels import _encode_cases
    _, host = gpu_libs
    _encode_cases(host, port, n, t, np.random.default_rng(n * 7 + t))


def test_encode_corrupt_decode_round_trip_on_device_encoder(gpu_libs):
    """200000 random data words -> bch_encode_batch -> 0..t flipped bits (none at bit 0) ->
    bch_decode_batch gives back the encoder's codewords with corrected = #flips"""
    _, host = gpu_libs
    n, t, B = 8191, 16, 200_000
    code = host.bch_gen_code(n, 2 * t + 1)
    k = code.data_bits
    dw = (k + 63) // 64
    data = synth.splitmix64(0xE1, B * dw).reshape(B, dw)
    clean = code.encode_batch(data, k)
    # every codeword of a linear code has zero syndromes: decoding the clean words changes nothing
    ok0, corr0, out0 = code.decode_batch(clean[:50000])
    assert (ok0 == 1).all() and (corr0 == 0).all() and (out0 == clean[:50000]).all()
    rng = np.random.default_rng(0xE1)
    words = clean.copy()
    nerr = rng.integers(0, t + 1, B)
    for e in range(1, t + 1):
        rows = np.nonzero(nerr >= e)[0]
        # distinct positions per word: stratify position e into its own band of the codeword
        band = (n - 1) // t
        pos = 1 + (e - 1) * band + rng.integers(0, band, rows.size)
        words[rows, pos // 64] ^= np.uint64(1) << (pos % 64).astype(np.uint64)
    ok, corr, out = code.decode_batch(words)
    assert (ok == 1).all() and (corr == nerr).all() and (out == clean).all()


@pytest.mark.parametrize("n,t", [(63, 5), (1023, 10), (8191, 16), (65535, 100)])
def test_closed_form_roots_of_low_degree_locators(gpu_libs, port, n, t):
    """k_bucket solves locators of degree 1 and 2 in closed form (bch.c:406-490 searches every point):
    same root lists as the oracle's search and as the device search with the shortcut switched off,
    including c1 = 0 (double root), c2 = 0 under L = 2 and roots outside the field"""
    cabi, _ = gpu_libs
    pc = port.code(n, t)
    h = _field(cabi, port, pc.m)
    c = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_code_create(h, 2 * t + 1, n, pc.gen.ctypes.data, pc.gen_degree,
                                             C.byref(c)), "code")
    d = 2 * t + 1
    rows = low_degree_locators(pc.m, t, np.random.default_rng(n))
    B = len(rows)
    elp = np.stack([e for _, e in rows])
    L = np.array([l for l, _ in rows], np.uint32)
    for direct in ("2", "0"):
        os.environ["BCHFFT_LOCATE_DIRECT"] = direct
        try:
            posn = np.zeros((B, d + 1), np.uint32)
            cnt = np.zeros(B, np.uint32)
            _lib.check(cabi, cabi.bchfft_bch_locate_host(c, elp.ctypes.data, L.ctypes.data, B,
                                                         posn.ctypes.data, cnt.ctypes.data), "locate")
        finally:
            del os.environ["BCHFFT_LOCATE_DIRECT"]
        for b in range(B):
            pr = pc.locate(elp[b, : L[b] + 1], int(L[b]))
            assert cnt[b] == len(pr) and (posn[b, : cnt[b]] == pr).all(), (direct, b, L[b], elp[b, :3])
    cabi.bchfft_code_destroy(c)
    cabi.bchfft_field_destroy(h)


def test_root_queue_drain_and_overflow_paths_on_device(gpu_libs):
    """k_locate_u collects roots in a shared-memory queue; BCHFFT_LOCU_QUEUE=6 forces both the
    mid-kernel drain and the queue-full direct path on the GPU.  Results must not change."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bch_fft_b200 import api, synth\n"
        "from oracle.pyoracle import Port\n"
        "host = api.Host(); P = Port(); n, t, B = 8191, 16, 6000\n"
        "code = host.bch_gen_code(n, 2 * t + 1); pc = P.code(n, t)\n"
        "bits = (synth.splitmix64(7, 8 * code.data_bits) & np.uint64(1)).astype(np.uint32).reshape(8, -1)\n"
        "pool = np.stack([code.encode(b) for b in bits])\n"
        "words, _ = synth.corrupt_codewords(pool, n, t, B, 7, beyond_fraction=0.05)\n"
        "ok, corr, out = code.decode_batch(words)\n"
        "rok, rcorr, rout = pc.decode_batch(words)\n"
        "assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()\n"
        "print('ok')\n")
    for q in ("6", "1"):
        env = dict(os.environ, BCHFFT_LOCU_QUEUE=q)
        res = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert res.returncode == 0 and "ok" in res.stdout, (q, res.stderr[-2000:])


@pytest.mark.parametrize("env", ["BCHFFT_BM_LPC=4", "BCHFFT_BM_LPC=16", "BCHFFT_BM_ALL_DISC=1", "BCHFFT_LOCU_NOSKIP=1",
                                 "BCHFFT_LOCATE_DIRECT=0", "BCHFFT_LOCATE_U=0"])
def test_alternative_kernel_routes_give_the_same_bits(gpu_libs, env):
    """every route switch of INTEGRATION.md section 6 that round 2 added: other lane-group sizes of the
    Berlekamp-Massey kernel, all discrepancies computed, zero forward values evaluated anyway, no
    closed-form roots, tile-engine root search -- same verdicts and words as the oracle at t = 16 and 64"""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bch_fft_b200 import api, synth\n"
        "from oracle.pyoracle import Port\n"
        "host = api.Host(); P = Port()\n"
        "for n, t, B in ((8191, 16, 5000), (8191, 64, 1500)):\n"
        "    code = host.bch_gen_code(n, 2 * t + 1); pc = P.code(n, t)\n"
        "    bits = (synth.splitmix64(11, 8 * code.data_bits) & np.uint64(1)).astype(np.uint32).reshape(8, -1)\n"
        "    pool = np.stack([code.encode(b) for b in bits])\n"
        "    words, _ = synth.corrupt_codewords(pool, n, t, B, 11, beyond_fraction=0.05)\n"
        "    ok, corr, out = code.decode_batch(words)\n"
        "    rok, rcorr, rout = pc.decode_batch(words)\n"
        "    assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all(), (n, t)\n"
        "print('ok')\n")
    k, v = env.split("=")
    res = subprocess.run([os.sys.executable, "-c", code], env=dict(os.environ, **{k: v}), capture_output=True,
                         text=True)
    assert res.returncode == 0 and "ok" in res.stdout, (env, res.stderr[-2000:])


def test_reference_demo_sources_build_and_pass(gpu_libs):
    """the reference's bchdemo source, UNCHANGED, linked against our libraries (drop-in check);
    needs the reference tree, so it only runs where /root/reference exists"""
    if not os.path.exists("/root/reference/bch/main.c"):
        pytest.skip("reference tree not present on this machine")
    subprocess.run(["make", "-C", ROOT, "demos"], check=True, stdout=subprocess.DEVNULL)
    for argv in (["8191", "16", "16"], ["65535", "100", "100"], ["8191", "200", "200"]):
        out = subprocess.run([os.path.join(ROOT, "host", "bchdemo")] + argv, capture_output=True, text=True)
        assert out.returncode == 0 and "Test successful" in out.stdout, out.stdout[-500:]


@pytest.mark.parametrize("m", [8, 12, 16])
def test_multiplicative_fft_vs_oracle_and_additive(gpu_libs, port, m):
    """BASELINE configs[4]: multiplicative FFT (2^16-1 = 3*5*17*257), full and truncated, against
    the oracle and against the additive path on the same polynomial (fft/test.c:147)"""
    cabi, host = gpu_libs
    n = 1 << m
    h = _field(cabi, port, m)
    B = 40
    polys = synth.random_polys(m, B, 0xC5)
    for nc in (n - 1, 0, 1, 2, 15, 255, (n - 1) // 3):
        nc = min(nc, n - 1)
        out = np.zeros((B, n), np.uint32)
        _lib.check(cabi, cabi.bchfft_multiplicative_fft_host(h, polys.ctypes.data, n, out.ctypes.data, n,
                                                             nc, B), "mfft")
        for b in (0, 33):
            want = port.mult_fft(m, polys[b, : n - 1]) if nc == n - 1 else port.mult_eval(m, polys[b], nc)
            assert (out[b, : n - 1] == want).all(), (m, nc, b)
    out = np.zeros((B, n), np.uint32)
    _lib.check(cabi, cabi.bchfft_multiplicative_fft_host(h, polys.ctypes.data, n, out.ctypes.data, n,
                                                         n - 1, B), "mfft")
    add, _ = _fft(cabi, h, polys, n - 1, natural=False)
    assert (out[:, : n - 1] == add).all()
    cabi.bchfft_field_destroy(h)
    for rec in [r for r in GOLDEN["multiplicative"] if r["m"] == m]:
        f = host.create_binary_finite_field(m)
        from tests.common import multiplicative_input
        poly = multiplicative_input(rec)
        if rec.get("truncated"):
            rc, got = f.evaluate_polynomial_multiplicative_FFT(poly, rec["num_coeffs"])
        else:
            data = poly.copy()
            rc, got = f.multiplicative_FFT(data), data[: n - 1]
        assert rc == 0 and f"{fnv64(got):016x}" == rec["fnv"]


def test_prebuilt_reference_demos_pass(gpu_libs):
    """host/fftdemo and host/bchdemo are the reference's own demo SOURCES (fft/test.c,
    bch/main.c) compiled unchanged against our libraries in the build container (`make demos`);
    here the binaries just have to run and report success"""
    fftdemo = os.path.join(ROOT, "host", "fftdemo")
    bchdemo = os.path.join(ROOT, "host", "bchdemo")
    if not (os.path.exists(fftdemo) and os.path.exists(bchdemo)):
        pytest.skip("demo binaries not built (need the reference tree at build time)")
    for argv in (["8191", "16", "16"], ["65535", "100", "100"], ["65535", "1000", "1000"], ["8191", "200", "200"]):
        out = subprocess.run([bchdemo] + argv, capture_output=True, text=True)
        assert out.returncode == 0 and "Test successful" in out.stdout, out.stdout[-500:]
    out = subprocess.run([fftdemo], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "Test successful!" in out.stdout, out.stdout[-800:]


def test_bchstream_cli_round_trip(gpu_libs, port, tmp_path):
    """host/bchstream (streaming decoder CLI) on the GPU: 100000 words of the
    (8191, 16) code through stdin/stdout in ragged batches; corrected words, ok and corrected equal
    bch_decode_batch's, and the first 2000 equal the oracle's"""
    exe = os.path.join(ROOT, "host", "bchstream")
    assert os.path.exists(exe), "host/bchstream missing: run `make stream` (or __graft_entry__.build())"
    _, host = gpu_libs
    n, t, B = 8191, 16, 100_000
    code = host.bch_gen_code(n, 2 * t + 1)
    k = code.data_bits
    dw = (k + 63) // 64
    words = code.encode_batch(synth.splitmix64(0x57, B * dw).reshape(B, dw), k)
    rng = np.random.default_rng(0x57)
    nerr = rng.integers(0, t + 3, B)                      # up to t+2 errors: failures included
    for e in range(1, t + 3):
        rows = np.nonzero(nerr >= e)[0]
        band = n // (t + 2)
        pos = (e - 1) * band + rng.integers(0, band, rows.size)   # bit 0 allowed (reference quirk)
        words[rows, pos // 64] ^= np.uint64(1) << (pos % 64).astype(np.uint64)
    ok, corr, out = code.decode_batch(words.copy())
    status = tmp_path / "status.bin"
    r = subprocess.run([exe, str(n), str(t), "-b", "30000", "-s", str(status)], input=words.tobytes(),
                       capture_output=True, check=True, timeout=600)
    got = np.frombuffer(r.stdout, np.uint64).reshape(B, -1)
    rec = np.fromfile(status, np.uint32).reshape(B, 2)
    assert (got == out).all() and (rec[:, 0] == ok).all() and (rec[:, 1] == corr).all()
    ok_w, corr_w, out_w = port.code(n, t).decode_batch(words[:2000].copy())
    assert (got[:2000] == out_w).all() and (rec[:2000, 0] == ok_w).all() and (rec[:2000, 1] == corr_w).all()


def test_host_api_concurrent_callers(gpu_libs, port):
    """re-entrancy of the drop-in library (threading row) on the real device: four
    threads with distinct buffers share one TbchCode / TfiniteField and mix bch_decode_batch,
    bch_decode and additive FFT calls; every answer equals the single-threaded one"""
    import threading
    _, host = gpu_libs
    n, t, B = 8191, 16, 3000
    code = host.bch_gen_code(n, 2 * t + 1)
    k = code.data_bits
    dw = (k + 63) // 64
    words = code.encode_batch(synth.splitmix64(0x7C, B * dw).reshape(B, dw), k)
    rng = np.random.default_rng(0x7C)
    nerr = rng.integers(0, t + 3, B)
    for e in range(1, t + 3):
        rows = np.nonzero(nerr >= e)[0]
        band = n // (t + 2)
        pos = (e - 1) * band + rng.integers(0, band, rows.size)
        words[rows, pos // 64] ^= np.uint64(1) << (pos % 64).astype(np.uint64)
    ok_w, corr_w, out_w = code.decode_batch(words)
    ok_o, corr_o, out_o = port.code(n, t).decode_batch(words[:300].copy())
    assert (ok_w[:300] == ok_o).all() and (corr_w[:300] == corr_o).all() and (out_w[:300] == out_o).all()
    f = host.create_binary_finite_field(13)
    poly = rng.integers(0, 1 << 13, 1 << 13, dtype=np.uint32)
    want_fft = port.additive_fft(13, poly, (1 << 13) - 1)[0]
    errors = []

    def worker(j):
        try:
            for rep in range(8):
                if (j + rep) % 3 == 0:
                    assert (f.evaluate_polynomial_additive_FFT(poly.copy(), (1 << 13) - 1) == want_fft).all()
                elif j % 2:
                    ok, corr, out = code.decode_batch(words)
                    assert (ok == ok_w).all() and (corr == corr_w).all() and (out == out_w).all()
                else:
                    for i in range(j, B, 211):
                        o, c, w = code.decode(words[i])
                        assert (o, c) == (int(ok_w[i]), int(corr_w[i])) and (w == out_w[i]).all()
        except Exception as e:   # noqa: BLE001 - reported to the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(j,)) for j in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("n,t,B", [(8191, 64, 12), (8191, 200, 7), (65535, 100, 6), (65535, 1000, 5)])
def test_wide_berlekamp_massey(gpu_libs, port, n, t, B):
    """k_bm_wide (one CTA per codeword; single decodes with large t) against k_bm and the oracle"""
    from tests.test_emu_kernels import _wide_bm_case
    cabi, host = gpu_libs
    _wide_bm_case(cabi, host, port, n, t, B, 1000 + t)
