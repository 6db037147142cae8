"""Per-kernel timing of ONE bch_decode call (host API) at n = 65535: where a single decode's
latency goes.  usage: python tools/gpu_single.py [t ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bch_fft_b200 import _lib, api, synth  # noqa: E402

lib = _lib.lib()
host = api.Host()
ulp = C.POINTER(C.c_ulong)
for t in [int(a) for a in sys.argv[1:]] or [100, 1000]:
    n = 65535
    code = host.bch_gen_code(n, 2 * t + 1)
    word = code.encode((synth.splitmix64(0xC2 + t, code.data_bits) & np.uint64(1)).astype(np.uint32))
    rng = np.random.default_rng(0xC2 + t)
    for p in rng.choice(np.arange(1, n), t, replace=False):
        word[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
    corr = C.c_uint32(0)

    def one():
        w = word.copy()
        return host.lib.bch_decode(C.byref(code.c), w.ctypes.data_as(ulp), None, C.byref(corr))

    for _ in range(3):
        one()
    t0 = time.perf_counter()
    for _ in range(10):
        one()
    wall = (time.perf_counter() - t0) / 10
    prof = _lib.profile(lib, one)
    print(f"t={t}: {1e3 * wall:.3f} ms per call; kernels:",
          {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])})
