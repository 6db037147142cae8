#!/usr/bin/env python
"""Small invocations of every kernel family, checked against the oracle; run under
compute-sanitizer (tools/gpu_sanitize.sh: memcheck / racecheck / synccheck / initcheck).
    python tools/sanitize_workload.py [light|full]
`light` keeps the racecheck pass (two orders of magnitude slower than native) to a few minutes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bch_fft_b200 import api, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.common import flip  # noqa: E402


def words_for(pc, n, t, B, rng):
    pool = [pc.encode(rng.integers(0, 2, n - pc.gen_degree, dtype=np.uint32)) for _ in range(4)]
    words = np.zeros((B, pc.words), np.uint64)
    for b in range(B):
        w = pool[b % 4].copy()
        ne = [0, 1, t, t // 2, t + 1, t + 2][b % 6] if b < 24 else int(rng.integers(0, t + 1))
        for p in set(int(x) for x in rng.choice(n, min(ne, n), replace=False)):
            flip(w, p)
        words[b] = w
    return words


def main():
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    host, port = api.Host(), pyoracle.Port()
    rng = np.random.default_rng(11)
    # decode: universal-LFSR syndromes + k_bm + k_locate_u (root queue) / general kernels / wide BM
    cases = [(8191, 16, 1500 if full else 200), (8191, 64, 40), (255, 8, 70), (65535, 16, 40)]
    if full:
        cases += [(8191, 200, 8), (65535, 1000, 2), (131071, 12, 33)]
    for n, t, B in cases:
        code, pc = host.bch_gen_code(n, 2 * t + 1), port.code(n, t)
        words = words_for(pc, n, t, B, rng)
        ok, corr, out = code.decode_batch(words)
        rok, rcorr, rout = pc.decode_batch(words)
        assert (ok == rok).all() and (corr == rcorr).all() and (out == rout).all(), (n, t)
        o, c, w = code.decode(words[2])
        assert (o, c) == (int(rok[2]), int(rcorr[2])) and (w == rout[2]).all()
        k = code.data_bits
        data = synth.splitmix64(n + t, 9 * ((k + 63) // 64)).reshape(9, -1)
        enc = code.encode_batch(data, k)
        bits = ((data[0][:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint32).reshape(-1)[:k]
        assert (enc[0] == pc.encode(bits)).all(), ("encode", n, t)
        print("decode/encode ok", n, t, B, flush=True)
    # additive FFT: single pass, plane-major (m = 16), multi-pass (m = 17 / 20), truncated
    for m, B in [(8, 40), (13, 40), (16, 33)] + ([(17, 3), (20, 2)] if full else []):
        n = 1 << m
        f = host.create_binary_finite_field(m)
        polys = synth.random_polys(m, B, 0x5A + m)
        for terms in (n - 1, 300 % n):
            res = f.additive_fft_batch(polys, terms)
            for b in (0, B - 1):
                assert (res[b] == port.additive_fft(m, polys[b], terms)[0]).all(), (m, terms, b)
        print("additive fft ok", m, B, flush=True)
    for m, B in [(8, 40), (12, 33)] + ([(16, 33)] if full else []):
        n = 1 << m
        f = host.create_binary_finite_field(m)
        polys = synth.random_polys(m, B, 0x5B + m)
        for nc in (n - 1, 15):
            res = f.multiplicative_fft_batch(polys, nc)
            want = port.mult_fft(m, polys[0, : n - 1]) if nc == n - 1 else port.mult_eval(m, polys[0], nc)
            assert (res[0] == want).all(), (m, nc)
        print("multiplicative fft ok", m, B, flush=True)
    host.shutdown()
    print("sanitize workload ok")


if __name__ == "__main__":
    main()
