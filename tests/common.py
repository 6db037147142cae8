"""Shared helpers: golden fixtures (tests/golden/golden_v1.json, produced from the unmodified
reference by tests/golden/make_golden.py) and the seeded inputs they were computed on."""
import json
import os

import numpy as np

from bch_fft_b200 import synth
from oracle.pyoracle import fnv64  # noqa: F401  (re-exported)

HERE = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(HERE, "golden", "golden_v1.json")) as _f:
    GOLDEN = json.load(_f)


def additive_input(rec):
    if rec.get("demo"):
        return synth.demo_poly(rec["m"])
    return synth.random_polys(rec["m"], 1, rec["seed"], zero_top=False)[0]


def multiplicative_input(rec):
    return synth.random_polys(rec["m"], 1, rec["seed"], zero_top=True)[0]


def decode_input(rec, encode):
    """rebuild the received words of a golden decode record; `encode(bits) -> codeword`"""
    n, t, seed, batch = rec["n"], rec["t"], rec["seed"], rec["batch"]
    gen_degree = next(c["gen_degree"] for c in GOLDEN["codes"] if c["n"] == n and c["t"] == t)
    bits = (synth.splitmix64(seed, 8 * (n - gen_degree)) & np.uint64(1)).astype(np.uint32).reshape(8, -1)
    pool = np.stack([encode(b) for b in bits])
    words, _ = synth.corrupt_codewords(pool, n, t, batch, seed, beyond_fraction=0.15)
    words[0] = pool[0]
    words[1] = pool[1]
    words[1, 0] ^= np.uint64(1)
    return pool, words


def flip(word, pos):
    word[pos >> 6] ^= np.uint64(1) << np.uint64(pos & 63)


def low_degree_locators(m, t, rng):
    """locators of degree <= 3 incl. the degenerate shapes of the closed-form path: c2 = 0 under L = 2,
    c1 = 0 (double root), c1 = c2 = 0, roots outside the field (Tr(c2 / c1^2) = 1)"""
    n, d = 1 << m, 2 * t + 1
    rows = []
    for L in (1, 2, 3):
        for _ in range(24):
            e = np.zeros(d + 1, np.uint32)
            e[0] = 1
            e[1: L + 1] = rng.integers(1, n, L)
            rows.append((L, e))
    for c1, c2 in ((0, 0), (0, 5), (7, 0), (0, 1), (1, 1), (1, 0), (n - 1, n - 1), (3, n - 2)):
        e = np.zeros(d + 1, np.uint32)
        e[0], e[1], e[2] = 1, c1, c2
        rows.append((2, e))
        rows.append((1, e))
    return rows
