"""Development aid: k_locate_u time per locator-degree bucket (all words of a batch get exactly e
errors).  Run on the GPU box: python tools/gpu_perk.py"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bch_fft_b200 import _lib, api, synth

lib = _lib.lib()
host = api.Host()
n, t, B = 8191, 16, 262144
code = host.bch_gen_code(n, 2 * t + 1)
bits = (synth.splitmix64(1, 8 * code.data_bits) & np.uint64(1)).astype(np.uint32).reshape(8, -1)
pool = np.stack([code.encode(b) for b in bits])
dfield = api.DeviceField(host.create_binary_finite_field(13), 0)
dcode = api.DeviceCode(dfield, code)
for e in (1, 3, 6, 12, 16):
    r = synth.splitmix64(e, B).astype(np.int64)
    words = pool[(r % 8)].copy()
    rng = np.random.default_rng(e)
    band = (n - 1) // t
    for k in range(e):
        pos = 1 + k * band + rng.integers(0, band, B)
        words[np.arange(B), pos // 64] ^= np.uint64(1) << (pos % 64).astype(np.uint64)
    d_words = torch.from_numpy(words.view(np.int64)).cuda()
    d_ok = torch.zeros(B, dtype=torch.uint8, device="cuda")
    d_corr = torch.zeros(B, dtype=torch.int32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    dcode.decode_dev(d_words.data_ptr(), B, d_ok.data_ptr(), d_corr.data_ptr(), sp)
    torch.cuda.synchronize()
    d_words.copy_(torch.from_numpy(words.view(np.int64)))
    prof = _lib.profile(lib, lambda: dcode.decode_dev(d_words.data_ptr(), B, d_ok.data_ptr(), d_corr.data_ptr(), sp))
    assert int(d_ok.sum()) == B and int(d_corr.sum()) == e * B
    print("errors", e, {k: round(v[0], 3) for k, v in prof.items()})
