"""Generates tests/golden/golden_v1.json from the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds (bch_fft_b200.synth), so only fingerprints and small outputs
are stored.  fnv64 = FNV-1a over the little-endian bytes of the array.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bch_fft_b200 import synth  # noqa: E402
from oracle.pyoracle import Ref, fnv64  # noqa: E402


def main():
    R = Ref()
    g = {"additive": [], "multiplicative": [], "codes": [], "decode": []}
    # additive FFT: random polynomials, several degree bounds
    for m, seed in ((8, 0xA8), (13, 0xAD), (16, 0xC1)):
        n = 1 << m
        poly = synth.random_polys(m, 1, seed, zero_top=False)[0]
        for terms in (0, 1, 2, 100 % n, 1000 % n, n // 2 - 1, n // 2, n - 2, n - 1):
            out, nat = R.additive_fft(m, poly, terms)
            rec = {"m": m, "seed": seed, "num_terms": terms, "fnv_result": f"{fnv64(out):016x}",
                   "fnv_natural": f"{fnv64(nat):016x}", "head": [int(x) for x in out[:8]]}
            if m == 8:
                rec["result"] = [int(x) for x in out]
            g["additive"].append(rec)
    for m in (16, 20):
        out, nat = R.additive_fft(m, synth.demo_poly(m), (1 << m) - 1)
        g["additive"].append({"m": m, "demo": True, "num_terms": (1 << m) - 1,
                              "fnv_result": f"{fnv64(out):016x}", "fnv_natural": f"{fnv64(nat):016x}",
                              "head": [int(x) for x in out[:8]],
                              "sum32": int(out.sum(dtype=np.uint64)) & 0xFFFFFFFF})
    # multiplicative FFT incl. truncated evaluation
    for m, seed in ((8, 0xB8), (12, 0xBC), (16, 0xC5)):
        n = 1 << m
        poly = synth.random_polys(m, 1, seed, zero_top=True)[0]
        full = R.mult_fft(m, poly[: n - 1])
        g["multiplicative"].append({"m": m, "seed": seed, "num_coeffs": n - 1,
                                    "fnv": f"{fnv64(full):016x}", "head": [int(x) for x in full[:8]]})
        for nc in (0, 1, 2, 15, 255 % n, (n - 1) // 3):
            ev = R.mult_eval(m, poly, nc)
            g["multiplicative"].append({"m": m, "seed": seed, "num_coeffs": nc, "truncated": True,
                                        "fnv": f"{fnv64(ev):016x}", "head": [int(x) for x in ev[:8]]})
    # codes: generator fingerprints
    for n, t in ((255, 8), (1023, 10), (8191, 8), (8191, 16), (8191, 64), (8191, 200), (65535, 16),
                 (65535, 100), (5000, 12)):
        c = R.code(n, t)
        g["codes"].append({"n": n, "t": t, "m": c.m, "gen_degree": c.gen_degree,
                           "fnv_gen": f"{fnv64(c.gen):016x}", "gen_word0": f"{int(c.gen[0]):016x}"})
    # decode: seeded batches (pool encoded by the reference itself)
    for n, t, seed, batch in ((255, 8, 0xD1, 64), (8191, 16, 0xC3, 96), (8191, 64, 0xC4, 48),
                              (8191, 200, 0xC5, 40), (65535, 100, 0xC2, 40), (5000, 12, 0xD2, 64)):
        c = R.code(n, t)
        bits = (synth.splitmix64(seed, 8 * (n - c.gen_degree)) & np.uint64(1)).astype(np.uint32)
        bits = bits.reshape(8, -1)
        pool = np.stack([c.encode(b) for b in bits])
        words, nerr = synth.corrupt_codewords(pool, n, t, batch, seed, beyond_fraction=0.15)
        words[0] = pool[0]                                  # clean word
        words[1] = pool[1]
        words[1, 0] ^= np.uint64(1)                         # single error at bit 0 (quirk)
        ok, corr, out = c.decode_batch(words)
        g["decode"].append({"n": n, "t": t, "seed": seed, "batch": batch,
                            "fnv_pool": f"{fnv64(pool):016x}", "fnv_in": f"{fnv64(words):016x}",
                            "ok": [int(x) for x in ok], "corrected": [int(x) for x in corr],
                            "fnv_out": f"{fnv64(out):016x}"})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden_v1.json:", {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
