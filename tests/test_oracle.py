"""The oracle is pinned here: the C restatement (oracle/bchfft_oracle.c) against
(1) the known answers the unmodified reference produced,
(2) the committed golden fixtures (tests/golden, generated from the reference), and
(3) oracle/_ref = the reference's own sources compiled in place, when it is present.
"""
import numpy as np
import pytest

from bch_fft_b200 import synth
from tests.common import GOLDEN, additive_input, decode_input, flip, fnv64, multiplicative_input


# ---- Appendix C known answers ---------------------------------------------------------------
def test_field_constants(port):
    for m, log3 in ((13, 934), (16, 61481), (20, 212012)):
        exp, log = port.tables(m)
        assert list(exp[1:5]) == [2, 4, 8, 16]
        assert log[3] == log3 and log[0] == (1 << m) - 1 and exp[(1 << m) - 1] == 0


def test_subspace_polynomial_constants(port):
    y13 = "1 6 168 1fa7 151a b5e 36b 50a 1cc1 97c 1d57 13e4 43".split()
    y16 = "1 6 168 47be b444 4c35 e9f1 73ad 368f d8f6 254a 4e55 d28f 5bfe 9458 6".split()
    for m, ys in ((13, y13), (16, y16)):
        coef, y = port.subspace_polys(m)
        assert [f"{v:x}" for v in y] == ys
        assert list(coef[1][:2]) == [1, 1]                       # P_1 = X^2 + X
        assert list(coef[2][:3]) == [6, 7, 1]                    # P_2 = X^4 + 7X^2 + 6X
        assert list(coef[3][:4]) == [0x770, 0x60c, 0x17d, 1]
    coef, _ = port.subspace_polys(13)
    assert [f"{v:x}" for v in coef[12][:13]] == \
        "1bb9 704 1a8c a49 19b4 41d 22b 1ba5 3b5 16f9 86b 1579 1".split()


@pytest.mark.parametrize("m,fnv,s32,head,fnv_nat", [
    (16, 0xda6143e12631d2df, 0x7fff8000, "3ff 4a7f 7e69 20a 624f c0b2 2e1 14f5", 0x4ce7d70654a732f3),
    (20, 0x43310189501ffec7, 0xfff80000, "3ff ea021 a8b08 7722a a19dd 4d471 79436 51cd", 0x15c5f83d417011fb),
])
def test_additive_demo_input(port, m, fnv, s32, head, fnv_nat):
    out, nat = port.additive_fft(m, synth.demo_poly(m), (1 << m) - 1)
    assert fnv64(out) == fnv and fnv64(nat) == fnv_nat
    assert int(out.sum(dtype=np.uint64)) & 0xFFFFFFFF == s32
    assert [f"{v:x}" for v in out[:8]] == head.split()


@pytest.mark.parametrize("n,t,deg,fnv,w0", [
    (8191, 8, 104, 0xf3f87fb3a08f8327, 0x0c138741c5c4fb23),
    (8191, 16, 208, 0x6f554676a1a13cac, 0xb378a601cdd0fdd1),
    (8191, 64, 832, 0xa5c9eb0bdfe5b7c8, 0x5ce9feb307fb770f),
    (8191, 200, 2444, 0x7f0c37ff95e5164e, 0x4a0250f8b2deabe7),
    (65535, 16, 256, 0x8d934af1db21a332, 0xa3d29ce86a53275b),
    (65535, 100, 1600, 0x51670241bd1287ea, 0xeb2de36fcbb50f4f),
    (65535, 1000, 15360, 0x11e8ab396e154f9c, 0x9d79553be50a3fa9),
    (1048575, 100, 2000, 0x38dc62bb051634f8, 0x3d091c15311e3ae5),
])
def test_generator_fingerprints(port, n, t, deg, fnv, w0):
    c = port.code(n, t)
    assert c.gen_degree == deg and fnv64(c.gen) == fnv and int(c.gen[0]) == w0


@pytest.mark.parametrize("n,t,errs,syn,L,elp,pos,corr,nonzero_words", [
    (8191, 16, (1, 100, 5000), "0 50e d0c 1be0 1b13 49 19cf 3b", 3, "1 50e 299 1102", (5000, 100, 1), 3, 0),
    (65535, 100, (3, 1000, 60000, 65534), "1 e8ee 7a51 711f 53af 9188 4822 8e72", 4,
     "1 e8ee 98a0 ff5 acdb", (65534, 60000, 1000, 3), 4, 0),
    (8191, 16, (0, 7), "0 81 37 1b01 515 bdd df8 1e10", 2, "1 81 80", (8191, 7), 2, 2),
])
def test_decoder_known_answers(port, n, t, errs, syn, L, elp, pos, corr, nonzero_words):
    c = port.code(n, t)
    w = np.zeros(c.words, np.uint64)
    for p in errs:
        flip(w, p)
    flag, s, _ = c.syndrome(w)
    # syndrome[0] is method-dependent: Appendix C lists the direct-path value
    assert [f"{v:x}" for v in s[1:8]] == syn.split()[1:]
    Lg, e = c.elp(s)
    assert Lg == L and [f"{v:x}" for v in e] == elp.split()
    assert tuple(c.locate(e, Lg)) == pos
    ok, cr, out = c.decode(w)
    assert (ok, cr) == (1, corr) and np.count_nonzero(out) == nonzero_words
    if nonzero_words:   # bit-0 quirk: bit 0 stays wrong, pad bit `order` is flipped
        assert out[0] == 1 and out[-1] == np.uint64(1) << np.uint64(63)


def test_decoder_behaviour_pins(port):
    # single error at bit 0: ok=1, corrected=1, reported position = order
    for n, t in ((8191, 8), (65535, 100)):
        c = port.code(n, t)
        w = np.zeros(c.words, np.uint64)
        w[0] = 1
        ok, cr, out = c.decode(w)
        assert (ok, cr) == (1, 1) and out[0] == 1
    # valid odd-weight codeword: flag 0 on the direct path, 1 (syndrome[0]=1) on the FFT path
    for n, t, want_flag in ((8191, 16, 0), (8191, 200, 1)):
        c = port.code(n, t)
        rng = np.random.default_rng(3)
        while True:
            cw = c.encode(rng.integers(0, 2, n - c.gen_degree, dtype=np.uint32))
            if sum(bin(int(x)).count("1") for x in cw) & 1:
                break
        flag, s, used_fft = c.syndrome(cw)
        assert flag == want_flag and used_fft == bool(want_flag) and s[0] == want_flag
        ok, cr, out = c.decode(cw)
        assert (ok, cr) == (1, 0) and (out == cw).all()


# ---- committed golden fixtures (generated from the unmodified reference) ----------------------
@pytest.mark.parametrize("rec", [r for r in GOLDEN["additive"] if r["m"] <= 16],
                         ids=lambda r: f"m{r['m']}-terms{r['num_terms']}")
def test_golden_additive(port, rec):
    out, nat = port.additive_fft(rec["m"], additive_input(rec), rec["num_terms"])
    assert f"{fnv64(out):016x}" == rec["fnv_result"]
    assert f"{fnv64(nat):016x}" == rec["fnv_natural"]
    assert [int(x) for x in out[:8]] == rec["head"]
    if "result" in rec:
        assert [int(x) for x in out] == rec["result"]


@pytest.mark.parametrize("rec", GOLDEN["multiplicative"],
                         ids=lambda r: f"m{r['m']}-nc{r['num_coeffs']}")
def test_golden_multiplicative(port, rec):
    m, n = rec["m"], 1 << rec["m"]
    poly = multiplicative_input(rec)
    got = port.mult_eval(m, poly, rec["num_coeffs"]) if rec.get("truncated") else port.mult_fft(m, poly[: n - 1])
    assert f"{fnv64(got):016x}" == rec["fnv"] and [int(x) for x in got[:8]] == rec["head"]


@pytest.mark.parametrize("rec", GOLDEN["codes"], ids=lambda r: f"n{r['n']}-t{r['t']}")
def test_golden_codes(port, rec):
    c = port.code(rec["n"], rec["t"])
    assert c.m == rec["m"] and c.gen_degree == rec["gen_degree"]
    assert f"{fnv64(c.gen):016x}" == rec["fnv_gen"] and f"{int(c.gen[0]):016x}" == rec["gen_word0"]


@pytest.mark.parametrize("rec", GOLDEN["decode"], ids=lambda r: f"n{r['n']}-t{r['t']}")
def test_golden_decode(port, rec):
    c = port.code(rec["n"], rec["t"])
    pool, words = decode_input(rec, c.encode)
    assert f"{fnv64(pool):016x}" == rec["fnv_pool"] and f"{fnv64(words):016x}" == rec["fnv_in"]
    ok, corr, out = c.decode_batch(words)
    assert [int(x) for x in ok] == rec["ok"] and [int(x) for x in corr] == rec["corrected"]
    assert f"{fnv64(out):016x}" == rec["fnv_out"]
    if rec["t"] <= 100:   # (at t=200 colliding error draws cancel and every word decodes)
        assert 0 in rec["ok"] and 1 in rec["ok"]      # the fixture covers both verdicts


# ---- live comparison with the compiled reference ----------------------------------------------
def test_port_vs_reference_additive(port, ref):
    rng = np.random.default_rng(1)
    for m in (4, 5, 8, 10, 13):
        n = 1 << m
        for deg in (0, 1, 2, 3, 7, 8, 100 % n, n // 2 - 1, n // 2, n - 2, n - 1):
            p = rng.integers(0, n, n, dtype=np.uint32)
            a, ac = port.additive_fft(m, p, deg)
            b, bc = ref.additive_fft(m, p, deg)
            assert (a == b).all() and (ac == bc).all(), (m, deg)
    p = rng.integers(0, 1 << 10, 1 << 10, dtype=np.uint32)
    nat = port.additive_fft(10, p, 1023)[1]
    for variant in ("additive_DFT_opt", "additive_DFT", "additive_DFT_ref"):
        assert (ref.additive_dft_variant(10, p, 1023, variant) == nat).all()


def test_port_vs_reference_multiplicative(port, ref):
    rng = np.random.default_rng(2)
    for m in (8, 10, 12, 14):
        n = 1 << m
        d = rng.integers(0, n, n - 1, dtype=np.uint32)
        assert (port.mult_fft(m, d) == ref.mult_fft(m, d)).all()
        p = np.zeros(n, np.uint32)
        p[: n - 1] = d
        assert (ref.additive_fft(m, p, n - 1)[0] == ref.mult_fft(m, d)).all()   # fft/test.c:147
        for nc in (0, 1, 2, 5, 15, 255 % n, n // 3):
            assert (port.mult_eval(m, p, nc) == ref.mult_eval(m, p, nc)).all(), (m, nc)


@pytest.mark.parametrize("n,t", [(255, 8), (8191, 16), (8191, 200), (65535, 16), (5000, 12)])
def test_port_vs_reference_bch(port, ref, n, t):
    rng = np.random.default_rng(7)
    pc, rc = port.code(n, t), ref.code(n, t)
    assert (pc.gen == rc.gen).all()
    bits = rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint32)
    cw = pc.encode(bits)
    assert (cw == rc.encode(bits)).all()
    for ne in (0, 1, t // 2, t, t + 1, t + 2):
        for trial in range(2):
            w = cw.copy()
            pos = rng.choice(n, ne, replace=False)
            if trial and ne:
                pos[0] = 0
            for p in pos:
                flip(w, int(p))
            fa, sa, ua = pc.syndrome(w)
            fb, sb, ub = rc.syndrome(w)
            assert fa == fb and (sa == sb).all() and ua == ub
            La, ea = pc.elp(sa)
            Lb, eb = rc.elp(sb)
            assert La == Lb and (ea == eb).all()
            if La <= t:
                assert (pc.locate(ea, La) == rc.locate(eb, Lb)).all()
            a, b = pc.decode(w), rc.decode(w)
            assert a[0] == b[0] and a[1] == b[1] and (a[2] == b[2]).all()
