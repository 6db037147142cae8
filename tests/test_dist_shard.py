"""Multi-rank path on the CPU: two gloo ranks each decode their contiguous shard of one batch
(the kernels run under the CPU emulator here), the only cross-rank traffic being the scalar
reductions bench.py does (max of times, sums of counters).  No data-path collective exists."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, emu_host, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    from bch_fft_b200 import api, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host = api.Host(emu_host)
    n, t, total = 255, 8, 200
    code = host.bch_gen_code(n, 2 * t + 1)
    bits = (synth.splitmix64(7, 8 * code.data_bits) & np.uint64(1)).astype(np.uint32).reshape(8, -1)
    pool = np.stack([code.encode(b) for b in bits])
    words, _ = synth.corrupt_codewords(pool, n, t, total, 7, beyond_fraction=0.1)
    lo, hi = synth.shard_range(total, rank, world)
    ok, corr, out = code.decode_batch(words[lo:hi])
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out)
    stats = torch.tensor([float(ok.sum()), float(corr.sum()), float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array(list(stats.numpy()) + [tmax.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decode(emu_libs, port, tmp_path):
    from bch_fft_b200 import synth
    from tests.conftest import EMU_HOST
    world = 2
    mp.spawn(_worker, args=(world, 29517, EMU_HOST, str(tmp_path)), nprocs=world, join=True)
    n, t, total = 255, 8, 200
    pc = port.code(n, t)
    bits = (synth.splitmix64(7, 8 * (n - pc.gen_degree)) & np.uint64(1)).astype(np.uint32).reshape(8, -1)
    pool = np.stack([pc.encode(b) for b in bits])
    words, _ = synth.corrupt_codewords(pool, n, t, total, 7, beyond_fraction=0.1)
    rok, rcorr, rout = pc.decode_batch(words)
    gathered = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert (gathered == rout).all()                       # host-side gather = concatenation
    stats = np.load(tmp_path / "stats.npy")
    assert stats[0] == rok.sum() and stats[1] == rcorr.sum() and stats[2] == total and stats[3] == world


def test_shard_ranges_cover_the_batch():
    from bch_fft_b200 import synth
    for total in (0, 1, 31, 32, 1000, 10**6):
        for world in (1, 2, 4, 8):
            spans = [synth.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
