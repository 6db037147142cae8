"""debug: per-sweep clock trace of CTA (0,0) (library built with -DBCHFFT_TRACE)"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bch_fft_b200 import _lib, api, synth
lib = _lib.lib(); host = api.Host()
raw = C.CDLL(_lib.CUDA_SO)
def dump(title, fn="bchfft_debug_trace"):
    buf = (C.c_ulonglong * 512)()
    n = getattr(raw, fn)(buf, 255)
    print("==", title, n, "marks")
    prev = None
    for k in range(n):
        tag, clk = buf[2*k], buf[2*k+1]
        if prev is not None and tag != 1 and tag != 10: print("   tag %4d  +%8d clk" % (tag, clk - prev))
        else: print("   tag %4d" % tag)
        prev = clk
# FFT m=16, 64 polys
f = host.create_binary_finite_field(16)
polys = synth.random_polys(16, 1024, 1)
dump("warm")
f.additive_fft_batch(polys, 65535); dump("discard")
code = host.bch_gen_code(8191, 33)
bits = (synth.splitmix64(1, 8*code.data_bits) & np.uint64(1)).astype(np.uint32).reshape(8,-1)
pool = np.stack([code.encode(b) for b in bits])
words,_ = synth.corrupt_codewords(pool, 8191, 16, 65536, 5)
code.decode_batch(words); dump("discard", "bchfft_debug_trace_bch")
code.decode_batch(words); dump("decode: k_locate CTA (0,0)", "bchfft_debug_trace_bch")
