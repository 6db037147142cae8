"""Deterministic synthetic workloads: the same byte buffers feed the CUDA
path, the oracle and the CPU baseline.  Self-contained 64-bit PRNG (splitmix64), never glibc
rand().  Pure numpy; no dependency on the oracle.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, count: int) -> np.ndarray:
    """count 64-bit outputs of splitmix64 started at `seed` (vectorised: output i depends only
    on seed + (i+1)*golden)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_polys(m: int, batch: int, seed: int, zero_top: bool = True) -> np.ndarray:
    """batch x 2^m uint32 coefficients uniform in [0, 2^m); coefficient n-1 set to 0 like
    fft/test.c:126 when zero_top."""
    n = 1 << m
    r = splitmix64(seed, batch * n)
    polys = (r >> np.uint64(17)).astype(np.uint32) & np.uint32(n - 1)
    polys = polys.reshape(batch, n)
    if zero_top:
        polys[:, n - 1] = 0
    return polys


def demo_poly(m: int) -> np.ndarray:
    """the input of the reference's additive FFT self-test (fft/test.c:121-126)"""
    n = 1 << m
    p = (np.arange(n, dtype=np.uint32) & np.uint32(0x3FF))
    p[n - 1] = 0
    return p


def corrupt_codewords(pool: np.ndarray, length: int, t: int, batch: int, seed: int,
                      beyond_fraction: float = 0.01, picks: int = 4):
    """Synthetic received words for a linear code.

    pool: [P, W] uint64 valid codewords (encoded by the caller).  Word i is the XOR of `picks`
    pool entries (still a codeword), then e_i bit errors at positions uniform in [0, length):
    e_i uniform in 0..t, except that a `beyond_fraction` of the words gets t+1..t+3 errors
    (uncorrectable path).  Positions are drawn independently, so two draws may coincide and
    cancel (the word then carries fewer errors).  Returns (words [batch, W] uint64,
    nerr [batch] int32 = number of draws).
    """
    P, W = pool.shape
    r = splitmix64(seed, batch * (picks + 2))
    r = r.reshape(batch, picks + 2)
    words = np.zeros((batch, W), np.uint64)
    for k in range(picks):
        words ^= pool[(r[:, k] % np.uint64(P)).astype(np.int64)]
    nerr = (r[:, picks] % np.uint64(t + 1)).astype(np.int32)
    if beyond_fraction > 0:
        sel = (r[:, picks + 1] % np.uint64(1 << 20)).astype(np.float64) / float(1 << 20) < beyond_fraction
        extra = ((r[:, picks + 1] >> np.uint64(24)) % np.uint64(3)).astype(np.int32) + 1
        nerr = np.where(sel, t + extra, nerr)
    maxe = int(nerr.max()) if batch else 0
    if maxe:
        pr = splitmix64(seed ^ 0x5DEECE66D, batch * maxe).reshape(batch, maxe)
        pos = (pr % np.uint64(length)).astype(np.int64)
        active = np.arange(maxe)[None, :] < nerr[:, None]
        rows = np.broadcast_to(np.arange(batch)[:, None], pos.shape)[active]
        p = pos[active]
        flat = words.reshape(-1)
        np.bitwise_xor.at(flat, rows * W + (p >> 6), np.uint64(1) << (p & 63).astype(np.uint64))
    return words, nerr


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous range [lo, hi) of `total` items owned by `rank`."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)
