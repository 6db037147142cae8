#!/usr/bin/env python
"""End-to-end throughput of the drop-in batch calls from ONE process on all visible GPUs (the host
library shards the caller's pinned array by contiguous ranges, one host thread per device), in both
result modes.  Prints one JSON line per (devices, mode).
    python tools/e2e_multi.py [codewords per device = 1000000] [polynomials per device = 2048]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(ndev, mode, per_dev, polys_per_dev):
    from bch_fft_b200 import _lib, api, synth
    cabi, host = _lib.lib(), api.Host()
    assert host.lib.bchfft_host_num_devices() == ndev
    n, t = 8191, 16
    code = host.bch_gen_code(n, 2 * t + 1)
    bits = (synth.splitmix64(0xC3, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([code.encode(b) for b in bits.reshape(64, -1)])
    base, _ = synth.corrupt_codewords(pool, n, t, per_dev, 0xC3, beyond_fraction=0.01)
    B, W = per_dev * ndev, code.words

    def pinned(nbytes):
        p = C.c_void_p()
        _lib.check(cabi, cabi.bchfft_pinned_alloc(nbytes, C.byref(p)), "pinned")
        return np.frombuffer((C.c_ubyte * nbytes).from_address(p.value), np.uint8), p

    hw, hp = pinned(B * W * 8)
    hw = hw.view(np.uint64).reshape(B, W)
    hok, okp = pinned(B)
    hcorr, corrp = pinned(B * 4)
    ulp = C.POINTER(C.c_ulong)
    ts = []
    for it in range(4):
        for d in range(ndev):
            hw[d * per_dev:(d + 1) * per_dev] = base
        t0 = time.perf_counter()
        rc = host.lib.bch_decode_batch(C.byref(code.c), C.cast(hp.value, ulp), B, C.cast(okp.value, C.POINTER(C.c_ubyte)),
                                       C.cast(corrp.value, C.POINTER(C.c_uint32)))
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    ok0 = hok[:per_dev].copy()
    for d in range(1, ndev):
        assert (hok[d * per_dev:(d + 1) * per_dev] == ok0).all()
        assert (hw[d * per_dev:(d + 1) * per_dev] == hw[:per_dev]).all()
    out = {"devices": ndev, "mode": mode, "decode_codewords_per_s": B / min(ts[1:]), "decode_ms": 1e3 * min(ts[1:]),
           "codewords": B, "ok_fraction": float(ok0.mean())}
    if polys_per_dev:
        m, nn = 16, 1 << 16
        f = host.create_binary_finite_field(m)
        Bf = polys_per_dev * ndev
        pin, pinp = pinned(Bf * nn * 4)
        pout, poutp = pinned(Bf * nn * 4)
        pin.view(np.uint32)[:] = np.tile(synth.random_polys(m, polys_per_dev, 0xC1).reshape(-1), ndev)
        u32p = C.POINTER(C.c_uint32)
        ts = []
        for it in range(3):
            t0 = time.perf_counter()
            rc = host.lib.additive_fft_batch(C.byref(f.c), C.cast(pinp.value, u32p), C.cast(poutp.value, u32p), nn - 1, Bf)
            ts.append(time.perf_counter() - t0)
            assert rc == 0
        out["fft_points_per_s"] = Bf * nn / min(ts[1:])
        out["fft_ms"] = 1e3 * min(ts[1:])
    print(json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
        return
    per_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    polys = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    from bch_fft_b200 import _lib
    total = _lib.lib().bchfft_device_count()
    g = 1
    while g <= total:
        for mode in ("positions", "words"):
            env = dict(os.environ, BCHFFT_DEVICES=",".join(str(i) for i in range(g)), BCHFFT_D2H_MODE=mode)
            env.pop("BCHFFT_DEVICE", None)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(g), mode, str(per_dev),
                            str(polys if mode == "positions" else 0)], env=env, check=False)
        g *= 2


if __name__ == "__main__":
    main()
