"""CPU-only execution of the REAL kernel sources (bch_fft_b200/csrc) by the fiber-based CUDA
emulator in tools/cudaemu, checked against the oracle and the golden fixtures.  This is test
infrastructure for a build container without a GPU: it validates index arithmetic, tiling
schedules and the bitsliced field arithmetic before GPU time is spent.  The `-m gpu` tests
repeat the same checks on the device through the nvcc-built library.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from bch_fft_b200 import _lib, api, synth
from tests.common import GOLDEN, additive_input, decode_input, flip, fnv64, low_degree_locators

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(cabi, port, m):
    exp, log = port.tables(m)
    h = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_field_create(0, m, exp.ctypes.data, log.ctypes.data, C.byref(h)), "field")
    return h


def _fft(cabi, h, polys, terms):
    B, n = polys.shape
    res = np.full((B, n), 0xDEADBEEF, np.uint32)
    nat = np.zeros((B, n), np.uint32)
    _lib.check(cabi, cabi.bchfft_additive_fft_host(h, polys.ctypes.data, n, res.ctypes.data, n,
                                                   nat.ctypes.data, n, terms, B), "fft")
    assert (res[:, n - 1] == 0xDEADBEEF).all()      # entry n-1 is never written (additive_fft.c:20)
    return res[:, : n - 1], nat


@pytest.mark.parametrize("m", [4, 5, 8, 10, 13])
def test_emu_additive_fft_vs_oracle(emu_libs, port, m):
    cabi, _ = emu_libs
    n = 1 << m
    h = _field(cabi, port, m)
    rng = np.random.default_rng(m)
    for B, terms in ((1, n - 1), (33, n - 1), (3, 0), (3, 1), (3, 2), (5, 5), (34, n // 2 - 1),
                     (2, n // 2), (2, n - 2), (2, n + 5)):
        if m >= 12:
            B = min(B, 2)
        polys = rng.integers(0, n, (B, n), dtype=np.uint32)
        res, nat = _fft(cabi, h, polys, terms)
        for b in range(B):
            o, c = port.additive_fft(m, polys[b], min(terms, n - 1))
            assert (res[b] == o).all() and (nat[b] == c).all(), (m, B, terms, b)
    cabi.bchfft_field_destroy(h)


def test_emu_additive_fft_gf2_17_vs_reference(emu_libs, ref):
    """a field above GF(2^16) with a single-depth Cantor chain (odd m): multi-pass schedule, 4-byte
    elements, 20-word plane stride; full and truncated degree against the unmodified reference"""
    _, host = emu_libs
    m, n = 17, 1 << 17
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, 2, 77 + m)
    for terms in (n - 1, 1000):
        want, _ = ref.additive_fft_batch_mt(m, polys, terms, 2)
        assert (f.additive_fft_batch(polys, terms) == want).all(), terms


@pytest.mark.parametrize("rec", [r for r in GOLDEN["additive"] if r["m"] <= 13],
                         ids=lambda r: f"m{r['m']}-terms{r['num_terms']}")
def test_emu_additive_fft_golden(emu_libs, port, rec):
    cabi, _ = emu_libs
    h = _field(cabi, port, rec["m"])
    res, nat = _fft(cabi, h, additive_input(rec)[None, :].copy(), rec["num_terms"])
    assert f"{fnv64(res[0]):016x}" == rec["fnv_result"] and f"{fnv64(nat[0]):016x}" == rec["fnv_natural"]
    cabi.bchfft_field_destroy(h)


def test_emu_multi_pass_schedules(port):
    """small tiles (BCHFFT_TMAX) force the multi-window pass schedule on small fields; run in a
    subprocess because the tile size is read when the field context is created"""
    code = (
        "import sys, ctypes as C, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bch_fft_b200 import _lib\n"
        "from oracle.pyoracle import Port\n"
        "P = Port(); cabi = _lib.bind(sys.argv[1]); rng = np.random.default_rng(9)\n"
        "for m in (5, 7, 8, 10):\n"
        "    n = 1 << m; exp, log = P.tables(m); h = C.c_void_p()\n"
        "    assert cabi.bchfft_field_create(0, m, exp.ctypes.data, log.ctypes.data, C.byref(h)) == 0\n"
        "    for B, terms in ((33, n - 1), (2, n // 2), (2, 3), (2, 0)):\n"
        "        polys = rng.integers(0, n, (B, n), dtype=np.uint32)\n"
        "        res = np.zeros((B, n), np.uint32)\n"
        "        assert cabi.bchfft_additive_fft_host(h, polys.ctypes.data, n, res.ctypes.data, n, None, n, terms, B) == 0\n"
        "        for b in range(B):\n"
        "            assert (res[b, :n-1] == P.additive_fft(m, polys[b], terms)[0]).all(), (m, terms)\n"
        "print('ok')\n")
    emu = os.path.join(ROOT, "bch_fft_b200", "csrc", "libbchfft_emu.so")
    # tmax >= 6 lets GF(2^8) take the plane-major path (Cantor basis conversion); BCHFFT_PLANE_T
    # = 8: one tile pass, 7 / 6: line pass(es) on the slab + tile passes
    for tmax, plane_t in (("2", "8"), ("3", "8"), ("4", "8"), ("6", "8"), ("7", "8"), ("6", "7"), ("6", "6")):
        env = dict(os.environ, BCHFFT_TMAX=tmax, BCHFFT_PLANE_T=plane_t)
        out = subprocess.run([sys.executable, "-c", code, emu], env=env, capture_output=True, text=True)
        assert out.returncode == 0 and "ok" in out.stdout, (tmax, out.stderr[-2000:])


def test_emu_plane_major_gf2_16(emu_libs, port):
    """GF(2^16), full degree: plane-major schedule (two one-plane passes of generalised Taylor
    levels, conversion tail folded into the first butterfly pass) against the oracle"""
    cabi, _ = emu_libs
    m, n = 16, 1 << 16
    h = _field(cabi, port, m)
    polys = np.random.default_rng(16).integers(0, n, (2, n), dtype=np.uint32)
    before = cabi.bchfft_launch_count()
    res, nat = _fft(cabi, h, polys, n - 1)
    assert cabi.bchfft_launch_count() - before == 7   # bitslice, 2 plane, 2 butterfly, 2 gathers
    for b in range(2):
        o, c = port.additive_fft(m, polys[b], n - 1)
        assert (res[b] == o).all() and (nat[b] == c).all()
    cabi.bchfft_field_destroy(h)


def _words_for(pc, n, t, B, rng):
    pool = [pc.encode(rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint32)) for _ in range(4)]
    words = np.zeros((B, pc.words), np.uint64)
    for b in range(B):
        w = pool[b % 4].copy()
        ne = [0, 1, 2, t, t // 2, t + 1, t + 2, t + 3][b % 8] if b < 56 else int(rng.integers(0, t + 1))
        pos = rng.choice(n, min(ne, n), replace=False)
        if b % 5 == 0 and ne:
            pos[0] = 0                       # bit-0 quirk (bch.c:479-488)
        for p in set(int(x) for x in pos):
            flip(w, p)
        words[b] = w
    return words


@pytest.mark.parametrize("n,t", [(15, 2), (63, 5), (255, 8), (255, 20), (1023, 10), (300, 3),
                                 (5000, 12), (8191, 16), (8191, 64), (65535, 16)])
def test_emu_bch_stages_and_decode_vs_oracle(emu_libs, port, n, t):
    cabi, _ = emu_libs
    pc = port.code(n, t)
    h = _field(cabi, port, pc.m)
    c = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_code_create(h, 2 * t + 1, n, pc.gen.ctypes.data, pc.gen_degree,
                                             C.byref(c)), "code")
    B = 70 if n < 2000 else 36
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
    assert (ok == 0).any() and (ok == 1).any()
    cabi.bchfft_code_destroy(c)
    cabi.bchfft_field_destroy(h)


@pytest.mark.parametrize("n,t", [(63, 5), (255, 8), (1023, 10), (8191, 16)])
def test_emu_closed_form_roots_of_low_degree_locators(emu_libs, port, n, t):
    """k_bucket solves locators of degree 1 and 2 in closed form (bch.c:406-490 searches every point):
    same root lists as the oracle's search, and as the device search with the shortcut switched off"""
    cabi, _ = emu_libs
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
    for direct in (("2", "0") if n < 2000 else ("2",)):     # the emulated search of GF(2^13) takes minutes
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


@pytest.mark.parametrize("lpc", [2, 4, 8, 16, 32])
@pytest.mark.parametrize("n,t", [(255, 20), (8191, 16), (8191, 64), (1023, 10)])
def test_emu_berlekamp_massey_lane_groups(emu_libs, port, monkeypatch, n, t, lpc):
    """k_bm with `lpc` lanes per codeword (the shape large-t codes take on the GPU so that the
    coefficient buffers stay in shared memory), forced through BCHFFT_BM_LPC: stage-level call
    (syndromes from the array) and fused decode (syndromes from the bitsliced accumulators)"""
    monkeypatch.setenv("BCHFFT_BM_LPC", str(lpc))
    monkeypatch.setenv("BCHFFT_BM_WIDE", "0")      # few codewords with t >= 24 would take k_bm_wide
    cabi, _ = emu_libs
    pc = port.code(n, t)
    h = _field(cabi, port, pc.m)
    c = C.c_void_p()
    _lib.check(cabi, cabi.bchfft_code_create(h, 2 * t + 1, n, pc.gen.ctypes.data, pc.gen_degree,
                                             C.byref(c)), "code")
    B, d = 37, 2 * t + 1
    words = _words_for(pc, n, t, B, np.random.default_rng(n + t + lpc))
    syn = np.zeros((B, d), np.uint32)
    flags = np.zeros(B, np.int32)
    _lib.check(cabi, cabi.bchfft_bch_syndrome_host(c, words.ctypes.data, B, syn.ctypes.data,
                                                   flags.ctypes.data), "syndrome")
    elp = np.zeros((B, d + 1), np.uint32)
    L = np.zeros(B, np.uint32)
    _lib.check(cabi, cabi.bchfft_bch_elp_host(c, syn.ctypes.data, B, elp.ctypes.data, L.ctypes.data), "elp")
    for b in range(B):
        Lr, er = pc.elp(syn[b])
        assert L[b] == Lr and (elp[b, : Lr + 1] == er).all(), (b, L[b], Lr)
    ok = np.zeros(B, np.uint8)
    corr = np.zeros(B, np.uint32)
    out = words.copy()
    _lib.check(cabi, cabi.bchfft_bch_decode_host(c, out.ctypes.data, B, ok.ctypes.data,
                                                 corr.ctypes.data), "decode")
    rok, rcorr, rw = pc.decode_batch(words)
    assert (ok == rok).all() and (corr == rcorr).all() and (out == rw).all()
    cabi.bchfft_code_destroy(c)
    cabi.bchfft_field_destroy(h)


@pytest.mark.parametrize("rec", [r for r in GOLDEN["decode"] if r["n"] <= 8191 and r["t"] <= 64],
                         ids=lambda r: f"n{r['n']}-t{r['t']}")
def test_emu_host_api_decode_golden(emu_libs, rec):
    """through the reference-shaped host API (bch_gen_code / bch_encode / bch_decode_batch)"""
    _, host = emu_libs
    code = host.bch_gen_code(rec["n"], 2 * rec["t"] + 1)
    grec = next(c for c in GOLDEN["codes"] if c["n"] == rec["n"] and c["t"] == rec["t"])
    assert code.gen_degree == grec["gen_degree"] and f"{fnv64(code.gen):016x}" == grec["fnv_gen"]
    pool, words = decode_input(rec, code.encode)
    assert f"{fnv64(pool):016x}" == rec["fnv_pool"] and f"{fnv64(words):016x}" == rec["fnv_in"]
    ok, corr, out = code.decode_batch(words)
    assert [int(x) for x in ok] == rec["ok"] and [int(x) for x in corr] == rec["corrected"]
    assert f"{fnv64(out):016x}" == rec["fnv_out"]
    # single-word entry point agrees with the batch
    for i in (0, 1, 2, len(words) - 1):
        o, c, w = code.decode(words[i])
        assert (o, c) == (rec["ok"][i], rec["corrected"][i]) and (w == out[i]).all()


def _encode_cases(host, port, n, t, rng):
    """bch_encode_batch against one bch_encode call per item (host library) and the oracle's encoder"""
    code = host.bch_gen_code(n, 2 * t + 1)
    pc = port.code(n, t)
    k = code.data_bits
    for nbits in sorted({k, k - 1, max(k - 67, 1), 1, min(64, k), min(65, k)}):
        B = 37
        bits = rng.integers(0, 2, (B, nbits), dtype=np.uint32)
        bits[0] = 0
        bits[1] = 1
        dw = (nbits + 63) // 64
        packed = np.zeros((B, dw), np.uint64)
        for i in range(nbits):
            packed[:, i // 64] |= bits[:, i].astype(np.uint64) << np.uint64(i % 64)
        packed[:, -1] |= np.uint64(0xFFFF) << np.uint64(48) if nbits % 64 and nbits % 64 <= 48 else np.uint64(0)
        out = code.encode_batch(packed, nbits)       # bits above data_bits in the last word are ignored
        for b in (0, 1, 2, B - 1):
            assert (out[b] == code.encode(bits[b])).all(), (n, t, nbits, b)
            assert (out[b] == pc.encode(bits[b])).all(), (n, t, nbits, b)
    with pytest.raises(Exception):
        code.encode_batch(np.zeros((2, (k + 64) // 64 + 1), np.uint64), k + 1)     # bch.c:495


@pytest.mark.parametrize("n,t", [(8191, 200), (1023, 10), (4000, 30)])
def test_emu_encode_batch_wide(emu_libs, port, n, t):
    """warp-per-codeword encoder (k_encode_wide): any generator degree >= 64 -- the (8191, 200) code
    (deg g = 2444) takes it by default, BCHFFT_ENC_WIDE=1 forces it for smaller codes.  Run in a
    subprocess because the code context (and its encoder choice) is cached per process."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bch_fft_b200 import api\n"
        "from oracle.pyoracle import Port\n"
        "from tests.conftest import EMU_HOST\n"
        "from tests.test_emu_kernels import _encode_cases\n"
        f"_encode_cases(api.Host(EMU_HOST), Port(), {n}, {t}, np.random.default_rng({n + t}))\n"
        "print('ok')\n")
    env = dict(os.environ, BCHFFT_ENC_WIDE="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


@pytest.mark.parametrize("n,t", [(255, 8), (8191, 16), (8191, 64), (5000, 12), (1023, 10)])
def test_emu_encode_batch(emu_libs, port, n, t):
    _, host = emu_libs
    _encode_cases(host, port, n, t, np.random.default_rng(n + t))


def test_emu_host_api_additive(emu_libs, port):
    _, host = emu_libs
    f = host.create_binary_finite_field(10)
    exp, log = port.tables(10)
    assert (f.exp == exp).all() and (f.log == log).all()
    rng = np.random.default_rng(4)
    poly = rng.integers(0, 1024, 1024, dtype=np.uint32)
    want, want_nat = port.additive_fft(10, poly, 700)
    p = poly.copy()
    got = f.evaluate_polynomial_additive_FFT(p, 700)
    assert (got == want).all() and (p == want_nat).all()     # p_poly clobbered like the reference
    for variant in ("additive_DFT_opt", "additive_DFT", "additive_DFT_ref"):
        p = poly.copy()
        assert (f.additive_DFT(p, 700, variant) == want_nat).all()
    polys = rng.integers(0, 1024, (5, 1024), dtype=np.uint32)
    res = f.additive_fft_batch(polys, 1023)
    for b in range(5):
        assert (res[b] == port.additive_fft(10, polys[b], 1023)[0]).all()


@pytest.mark.parametrize("m", [8, 9, 10, 12])
def test_emu_multiplicative_fft_vs_oracle(emu_libs, port, m):
    """secondary kernel: mixed-radix FFT over the cyclic group, full and truncated"""
    cabi, host = emu_libs
    n = 1 << m
    h = _field(cabi, port, m)
    rng = np.random.default_rng(50 + m)
    B = 34
    polys = rng.integers(0, n, (B, n), dtype=np.uint32)
    for nc in (n - 1, 0, 1, 2, 5, 15, (n - 1) // 3, (n - 1) // 3 + 1):
        out = np.zeros((B, n), np.uint32)
        _lib.check(cabi, cabi.bchfft_multiplicative_fft_host(h, polys.ctypes.data, n, out.ctypes.data, n,
                                                             nc, B), "mfft")
        for b in (0, 31, 33):
            want = port.mult_fft(m, polys[b, : n - 1]) if nc == n - 1 else port.mult_eval(m, polys[b], nc)
            assert (out[b, : n - 1] == want).all(), (m, nc, b)
    cabi.bchfft_field_destroy(h)
    # additive and multiplicative paths agree (the relation fft/test.c:147 checks)
    f = host.create_binary_finite_field(m)
    p = polys[0].copy()
    p[n - 1] = 0
    mult = p.copy()
    assert f.multiplicative_FFT(mult) == 0
    add = f.evaluate_polynomial_additive_FFT(p.copy(), n - 1)
    assert (mult[: n - 1] == add).all()
    rc, ev = f.evaluate_polynomial_multiplicative_FFT(p, 7)
    assert rc == 0 and (ev == port.mult_eval(m, p, 7)).all()


def test_emu_multiplicative_fft_unknown_factorisation(emu_libs):
    _, host = emu_libs
    f = host.create_binary_finite_field(5)
    data = np.arange(32, dtype=np.uint32)
    assert f.multiplicative_FFT(data) == 1          # multiplicative_fft.c:110-114


def test_emu_edge_cases_and_errors(emu_libs, port):
    """empty and one-item batches, argument errors, unsupported field degrees (C-ABI contract)"""
    cabi, host = emu_libs
    h = _field(cabi, port, 8)
    n = 256
    polys = np.arange(n, dtype=np.uint32)[None, :].copy()
    res = np.full((1, n), 7, np.uint32)
    # empty batch: ok, nothing written
    assert cabi.bchfft_additive_fft_host(h, polys.ctypes.data, n, res.ctypes.data, n, None, n, n - 1, 0) == 0
    assert (res == 7).all()
    # null / short-stride arguments are rejected with a message
    assert cabi.bchfft_additive_fft_host(h, None, n, res.ctypes.data, n, None, n, n - 1, 1) < 0
    assert cabi.bchfft_additive_fft_dev(h, polys.ctypes.data, n - 1, res.ctypes.data, n, None, n, n - 1, 1, None) < 0
    assert b"stride" in cabi.bchfft_last_error()
    # one item
    assert cabi.bchfft_additive_fft_host(h, polys.ctypes.data, n, res.ctypes.data, n, None, n, n - 1, 1) == 0
    assert (res[0, : n - 1] == port.additive_fft(8, polys[0], n - 1)[0]).all()
    # field tables that are not the reference's are refused
    exp, log = port.tables(8)
    bad = exp.copy()
    bad[8] ^= 2
    hh = C.c_void_p()
    assert cabi.bchfft_field_create(0, 8, bad.ctypes.data, log.ctypes.data, C.byref(hh)) < 0
    e26 = np.zeros(4, np.uint32)
    assert cabi.bchfft_field_create(0, 26, e26.ctypes.data, e26.ctypes.data, C.byref(hh)) == -3
    # code creation argument checks (bch.c:37: distance >= length is no code)
    c = C.c_void_p()
    assert cabi.bchfft_code_create(h, 255, 255, None, 0, C.byref(c)) < 0
    assert cabi.bchfft_code_create(h, 2, 255, None, 0, C.byref(c)) < 0          # nothing to correct
    assert cabi.bchfft_code_create(h, 4, 255, None, 0, C.byref(c)) == 0         # even distances are codes (bch.c:37)
    cabi.bchfft_code_destroy(c)
    pc = port.code(255, 8)
    assert cabi.bchfft_code_create(h, 17, 255, pc.gen.ctypes.data, pc.gen_degree, C.byref(c)) == 0
    words = np.zeros((1, pc.words), np.uint64)
    ok = np.zeros(1, np.uint8)
    corr = np.full(1, 9, np.uint32)
    assert cabi.bchfft_bch_decode_host(c, words.ctypes.data, 0, ok.ctypes.data, corr.ctypes.data) == 0
    assert corr[0] == 9
    flip(words[0], 200)
    assert cabi.bchfft_bch_decode_host(c, words.ctypes.data, 1, ok.ctypes.data, corr.ctypes.data) == 0
    assert ok[0] == 1 and corr[0] == 1 and not words.any()
    cabi.bchfft_code_destroy(c)
    cabi.bchfft_field_destroy(h)
    # host API: the reference's failure conventions
    with pytest.raises(ValueError):
        host.bch_gen_code(255, 255)                       # m_packed_gen_poly == NULL
    with pytest.raises(ValueError):
        host.create_binary_finite_field(31)               # exp == log == NULL


@pytest.mark.parametrize("m", [8, 9, 12])
def test_host_only_multiplicative_building_blocks(emu_libs, port, m):
    """multiplicative_FFT_step_chaining / _compute_step / _reorder_step (host-only symbols the
    reference's libfft exports, libfft/multiplicative_fft.c:124-257) against the oracle's
    transform and the naive DFT_direct; the log-form input survives a compute step."""
    import ctypes as C
    _, host = emu_libs
    L = host.lib
    f = host.create_binary_finite_field(m)
    n, order = f.n, f.n - 1
    u32p = C.POINTER(C.c_uint32)
    L.numFactors.restype = C.c_uint32
    L.numFactors.argtypes = [C.c_uint32]
    L.factors.restype = u32p
    L.factors.argtypes = [C.c_uint32]
    nf, facts = L.numFactors(m), L.factors(m)
    rng = np.random.default_rng(900 + m)
    data = rng.integers(0, n, order, dtype=np.uint32)
    data[rng.integers(0, order, order // 8)] = 0          # logzero entries take the skip branch
    want = port.mult_fft(m, data)
    logs = np.ctypeslib.as_array(f.c.log, (n,))[data].astype(np.uint32)
    buf = np.zeros(order, np.uint32)
    fp = C.byref(f.c)
    as_p = lambda a: a.ctypes.data_as(u32p)
    # one compute step on its own: input logs unchanged, rows are size-N DFTs
    before = logs.copy()
    local = order // facts[0]
    L.multiplicative_FFT_compute_step(fp, as_p(logs), as_p(buf), C.c_uint32(local), C.c_uint32(1), C.c_uint32(1))
    assert (logs == before).all()
    N, count = order // local, local
    exp = np.ctypeslib.as_array(f.c.exp, (n,))
    for i, s in ((0, 0), (1, 0), (N - 1, count - 1)):
        acc = 0
        for j in range(N):
            d = int(before[s + count * j])
            if d != f.c.logzero:
                acc ^= int(exp[(d + local * i * j) % order])
        assert int(buf[s + count * i]) == acc
    # the whole chain
    L.multiplicative_FFT_step_chaining(fp, as_p(logs), as_p(buf), facts, C.c_uint32(nf), C.c_uint32(order),
                                       C.c_uint32(1))
    assert (logs == want).all()
    if m <= 9:
        naive = data.copy()
        L.DFT_direct(fp, as_p(naive), as_p(buf))
        assert (naive == want).all()


@pytest.mark.parametrize("m", [4, 8, 11])
def test_host_only_check_fft_polynomials(emu_libs, port, m):
    """check_fft_polynomials (libfft/additive_fft.c:181-262) accepts the subspace polynomials
    P_i = P_(i-1)^2 + y_(i-1) P_(i-1) and rejects a corrupted table"""
    import ctypes as C
    _, host = emu_libs
    L = host.lib
    f = host.create_binary_finite_field(m)
    u32p = C.POINTER(C.c_uint32)
    L.check_fft_polynomials.restype = C.c_int
    mul = lambda a, b: int(f.multiply(a, b))
    div = np.zeros((m, m + 1), np.uint32)
    div[0, 1] = 1                                          # P_0 = X
    for i in range(1, m):
        prev = div[i - 1]
        y = 0                                              # P_(i-1)(2^(i-1))
        for j in range(i + 1):
            v = 1 << (i - 1)
            for _ in range(j - 1):
                v = mul(v, v)
            if j >= 1:
                y ^= mul(int(prev[j]), v)
        for j in range(1, i + 2):
            sq = mul(int(prev[j - 1]), int(prev[j - 1])) if j >= 2 else 0
            div[i, j] = sq ^ mul(y, int(prev[j]))
    degrees = np.array([0] + [1 << j for j in range(m + 1)], np.uint32)
    yv = np.zeros(m, np.uint32)
    args = lambda d: (C.byref(f.c), d.ctypes.data_as(u32p), degrees.ctypes.data_as(u32p), yv.ctypes.data_as(u32p))
    assert L.check_fft_polynomials(*args(div)) == 1
    bad = div.copy()
    bad[m - 1, 1] ^= 1
    assert L.check_fft_polynomials(*args(bad)) == 0


def test_emu_bchstream_cli(emu_libs, port, tmp_path):
    """host/bchstream: packed codewords stdin -> stdout through bch_decode_batch,
    several batches incl. a ragged last one; output, ok and corrected equal the oracle's"""
    import subprocess
    subprocess.run(["make", "-C", ROOT, "stream-emu"], check=True, stdout=subprocess.DEVNULL)
    n, t = 255, 4
    pc = port.code(n, t)
    rng = np.random.default_rng(77)
    B, W = 70, 4
    words = np.zeros((B, W), np.uint64)
    for i in range(B):
        bits = rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint8)
        w = pc.encode(bits)
        ne = int(rng.integers(0, t + 3))
        for p in rng.choice(n, ne, replace=False):
            w[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
        words[i] = w
    ok_w, corr_w, out_w = pc.decode_batch(words.copy())
    status = tmp_path / "status.bin"
    exe = os.path.join(ROOT, "host", "bchstream_emu")
    r = subprocess.run([exe, str(n), str(t), "-b", "32", "-s", str(status)], input=words.tobytes(),
                       capture_output=True, check=True)
    got = np.frombuffer(r.stdout, np.uint64).reshape(B, W)
    assert (got == out_w).all()
    rec = np.fromfile(status, np.uint32).reshape(B, 2)
    assert (rec[:, 0] == ok_w).all() and (rec[:, 1] == corr_w).all()
    assert b"70 codewords" in r.stderr
    # truncated input: the stray bytes are reported, the exit code is non-zero
    r = subprocess.run([exe, str(n), str(t)], input=words.tobytes()[:-3], capture_output=True)
    assert r.returncode != 0 and len(r.stdout) == (B - 1) * W * 8


def test_emu_host_api_concurrent_callers(emu_libs, port):
    """the reference is re-entrant (threading row): concurrent callers with distinct
    buffers and shared read-only TfiniteField / TbchCode must get the single-threaded answers.
    Four threads mix bch_decode_batch, bch_decode and additive FFT calls on two codes."""
    import threading
    _, host = emu_libs
    rng = np.random.default_rng(4242)
    jobs = []
    for n, t in ((255, 4), (1023, 6)):
        code = host.bch_gen_code(n, 2 * t + 1)
        pc = port.code(n, t)
        W = code.words
        words = np.zeros((40, W), np.uint64)
        for i in range(40):
            w = pc.encode(rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint8))
            for p in rng.choice(n, int(rng.integers(0, t + 2)), replace=False):
                w[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
            words[i] = w
        jobs.append((code, words, pc.decode_batch(words.copy())))
    f = host.create_binary_finite_field(10)
    poly = rng.integers(0, 1 << 10, 1 << 10, dtype=np.uint32)
    want_fft = port.additive_fft(10, poly, (1 << 10) - 1)[0]
    errors = []

    def worker(k):
        try:
            for rep in range(6):
                code, words, (ok_w, corr_w, out_w) = jobs[(k + rep) % 2]
                if (k + rep) % 3 == 0:
                    got = f.evaluate_polynomial_additive_FFT(poly.copy(), (1 << 10) - 1)
                    assert (got == want_fft).all()
                elif k % 2:
                    ok, corr, out = code.decode_batch(words)
                    assert (ok == ok_w).all() and (corr == corr_w).all() and (out == out_w).all()
                else:
                    for i in range(0, 40, 7):
                        o, c, w = code.decode(words[i])
                        assert (o, c) == (int(ok_w[i]), int(corr_w[i])) and (w == out_w[i]).all()
        except Exception as e:   # noqa: BLE001 - reported to the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def _wide_bm_case(cabi, host, port, n, t, B, seed):
    """few codewords, large t: launch_bm picks k_bm_wide (one CTA per codeword); same answers as
    the one-thread-per-codeword kernel (BCHFFT_BM_WIDE=0) and as the oracle, >t errors included"""
    code = host.bch_gen_code(n, 2 * t + 1)
    pc = port.code(n, t)
    rng = np.random.default_rng(seed)
    words = np.zeros((B, code.words), np.uint64)
    for i in range(B):
        w = pc.encode(rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint8))
        ne = [0, 1, t, t + 1, t + 7][i] if i < 5 else int(rng.integers(0, t + 3))
        for p in rng.choice(n, ne, replace=False):
            w[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
        words[i] = w
    want = pc.decode_batch(words.copy())
    names = {}

    def run():
        names["got"] = code.decode_batch(words)
    prof = _lib.profile(cabi, run)
    assert any(k.startswith("k_bm_wide") for k in prof), sorted(prof)
    # t >= 48 (nodd >= 48) and at most 64 words: syndromes through one additive FFT per word
    assert any(k.startswith("k_syn_from_fft") for k in prof) == (t >= 48), sorted(prof)
    for a, b in zip(names["got"], want):
        assert (a == b).all()
    os.environ["BCHFFT_BM_WIDE"] = "0"
    os.environ["BCHFFT_SYN_FFT"] = "0"
    try:
        prof = _lib.profile(cabi, run)
    finally:
        del os.environ["BCHFFT_BM_WIDE"]
        del os.environ["BCHFFT_SYN_FFT"]
    assert not any(k.startswith("k_bm_wide") for k in prof) and any(k.startswith("k_bm") for k in prof)
    assert not any(k.startswith("k_syn_from_fft") for k in prof)
    for a, b in zip(names["got"], want):
        assert (a == b).all()
    # stage-level entry point (bch_compute_elp): L and the whole locator buffer
    for i in (0, 1, 2, 3, B - 1):
        flag_w, syn, _ = pc.syndrome(words[i])
        flag_g, syn_g = code.eval_syndrome(words[i])          # bch_eval_syndrome incl. entry 0 and the flag
        assert flag_g == flag_w and (syn_g == syn).all()
        L_w, elp_w = pc.elp(syn)
        L_g, elp_g = code.compute_elp(syn)
        assert L_g == L_w and (np.asarray(elp_g)[: L_w + 1] == elp_w).all()


def test_emu_wide_berlekamp_massey(emu_libs, port):
    cabi, host = emu_libs
    _wide_bm_case(cabi, host, port, 8191, 64, 9, 640)


@pytest.mark.parametrize("n,d", [(255, 10), (1023, 6), (600, 12), (8191, 4)])
def test_emu_even_designed_distance(emu_libs, port, ref, n, d):
    """the reference accepts any designed distance d < length (bch.c:37); an even d corrects
    t = (d-1)/2 errors and runs d-1 Berlekamp-Massey iterations.  Stages and decode against the
    restatement and the unmodified reference."""
    cabi, host = emu_libs
    t = (d - 1) // 2
    code = host.bch_gen_code(n, d)
    pc, rc = port.code(n, t, d=d), ref.code(n, t, d=d)
    assert code.gen_degree == rc.gen_degree and (code.gen == rc.gen).all() and (pc.gen == rc.gen).all()
    B = 40
    words = _words_for(pc, n, t, B, np.random.default_rng(n + d))
    ok, corr, out = code.decode_batch(words)
    for chk in (pc, rc):
        rok, rcorr, rw = chk.decode_batch(words)
        assert (ok == rok).all() and (corr == rcorr).all() and (out == rw).all(), chk
    for b in range(0, B, 3):
        fl, syn = code.eval_syndrome(words[b])
        rfl, rsyn, _ = rc.syndrome(words[b])
        assert fl == rfl and (syn == rsyn).all()
        L, elp = code.compute_elp(syn)
        rL, relp = rc.elp(syn)
        assert L == rL and (elp == relp).all()


def test_emu_gen_code_rejects_what_the_device_cannot_decode(emu_libs):
    _, host = emu_libs
    for n, d in ((255, 2), (255, 1), ((1 << 26) - 1, 5)):
        with pytest.raises(ValueError):
            host.bch_gen_code(n, d)


def test_emu_decode_result_modes_agree(emu_libs, port):
    """bch_decode_batch returns results in one of three ways (BCHFFT_D2H_MODE): position lists applied by
    host threads, every word copied back, or the changed words patched into the caller's array by the
    device; all must give the reference's words, untouched failures included"""
    _, host = emu_libs
    n, t, B = 1023, 10, 300
    code, pc = host.bch_gen_code(n, 2 * t + 1), port.code(n, t)
    words = _words_for(pc, n, t, B, np.random.default_rng(99))
    rok, rcorr, rout = pc.decode_batch(words)
    try:
        for mode in ("positions", "patch", "words"):
            os.environ["BCHFFT_D2H_MODE"] = mode
            ok, corr, out = code.decode_batch(words)
            assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all(), mode
        # five chunks through three lanes: the staging of a lane is reused while host threads still flip
        os.environ["BCHFFT_D2H_MODE"] = "positions"
        os.environ["BCHFFT_E2E_CHUNK"] = "64"
        os.environ["BCHFFT_FLIP_THREADS"] = "3"
        ok, corr, out = code.decode_batch(words)
        assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()
    finally:
        for k in ("BCHFFT_D2H_MODE", "BCHFFT_E2E_CHUNK", "BCHFFT_FLIP_THREADS"):
            os.environ.pop(k, None)


def test_emu_in_library_multi_device_sharding(emu_libs, port):
    """host library with a set of devices: batch calls cut the caller's arrays into one
    contiguous range per device, one host thread each; results equal the single-device ones.  The
    emulator pretends to have three devices (BCHFFT_EMU_DEVICES); small BCHFFT_MIN_PER_DEVICE forces the
    split on a small batch."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from bch_fft_b200 import api, synth\n"
        "from oracle.pyoracle import Port\n"
        "from tests.conftest import EMU_HOST\n"
        "from tests.test_emu_kernels import _words_for\n"
        "host, P = api.Host(EMU_HOST), Port()\n"
        "assert host.lib.bchfft_host_num_devices() == 3\n"
        "n, t, B = 255, 8, 1000\n"
        "code, pc = host.bch_gen_code(n, 2 * t + 1), P.code(n, t)\n"
        "words = _words_for(pc, n, t, B, np.random.default_rng(5))\n"
        "ok, corr, out = code.decode_batch(words)\n"
        "rok, rcorr, rout = pc.decode_batch(words)\n"
        "assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all()\n"
        "k = code.data_bits; dw = (k + 63) // 64\n"
        "data = synth.splitmix64(3, B * dw).reshape(B, dw)\n"
        "enc = code.encode_batch(data, k)\n"
        "bits = lambda i: ((data[i][:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint32).reshape(-1)[:k]\n"
        "for i in (0, 399, 400, 999): assert (enc[i] == pc.encode(bits(i))).all()\n"
        "m = 8; f = host.create_binary_finite_field(m)\n"
        "polys = synth.random_polys(m, 200, 9)\n"
        "res = f.additive_fft_batch(polys, 255)\n"
        "mul = f.multiplicative_fft_batch(polys, 255)\n"
        "for b in (0, 63, 64, 127, 128, 199):\n"
        "    assert (res[b] == P.additive_fft(m, polys[b], 255)[0]).all() and (mul[b] == res[b]).all()\n"
        "print('contexts', host.lib.bchfft_host_num_contexts())\n")
    env = dict(os.environ, BCHFFT_EMU_DEVICES="3", BCHFFT_MIN_PER_DEVICE="64")
    env.pop("BCHFFT_DEVICE", None)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    # one field (GF(2^8)) + one code context per device
    assert "contexts 6" in res.stdout, res.stdout


def test_emu_multiplicative_fft_gf2_16_fast_path(emu_libs, port, ref):
    """GF(2^16): Good-Thomas 255 x 257, cyclotomic 257-point transforms, 3 x 5 x 17 prime-factor transforms
    (mfft16_kernels.cuh) against the restatement and the unmodified reference, full length and truncated
    coefficient counts on both sides of the fast path's threshold (255)"""
    _, host = emu_libs
    m, n = 16, 1 << 16
    f = host.create_binary_finite_field(m)
    polys = synth.random_polys(m, 33, 0xC5)
    full = f.multiplicative_fft_batch(polys, n - 1)
    assert (full[0] == port.mult_fft(m, polys[0, : n - 1])).all()
    assert (full[32] == ref.mult_fft(m, polys[32, : n - 1])).all()
    for nc in (255, 256, 40000):
        got = f.multiplicative_fft_batch(polys, nc)
        assert (got[32] == ref.mult_eval(m, polys[32], nc)).all(), nc
