#!/usr/bin/env python
"""Benchmark of the hot path on every BASELINE.json configuration, in one JSON line.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU code, all cores

Headline (`metric`, `value`, `roofline`, `e2e`, `cpu_baseline`): BASELINE configs[2], batched BCH decode
over GF(2^13), n = 8191, t = 16, 10^6 synthetic codewords per GPU with injected errors (configs[1] is a
single decode and carries no throughput).  A "step" is one pass of the hot path over one batch of
synthetic input resident in HBM.  The other configurations ride in the same line, each with its own
value / roofline / cpu_baseline / e2e / parity:

    fft                                BASELINE configs[0] batched: additive FFT GF(2^16), 4096 polynomials (C1)
    configs.single_call                configs[0] / configs[1] as ONE call of the reference's entry points (C1, C2)
    configs.decode_n8191_t64 / _t200   configs[2] on the survey's two other codes (C3)
    configs.additive_fft_gf2_20        configs[3]: additive FFT GF(2^20), 256 polynomials (C4)
    configs.multiplicative_fft_gf2_16  configs[4]: multiplicative FFT GF(2^16), 4096 polynomials (C5)
    configs.encode_n8191_t16           the systematic encoder next to the decoder

Multi-GPU: one process per GPU (torchrun), each rank works on its own range of codewords / polynomials,
no collective on the data path ("weak" scaling: per-GPU batch fixed).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bch_fft_b200 import synth  # noqa: E402

METRIC = "bch_codewords_decoded_per_s"
SM_COUNT, ALU_LANES_PER_SM_CLK = 148, 64      # B200: 4 SMSPs x 16 lanes per clock on the alu pipe


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--length", type=int, default=8191)
    ap.add_argument("--t", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1_000_000, help="codewords per GPU (headline)")
    ap.add_argument("--fft-m", type=int, default=16)
    ap.add_argument("--fft-batch", type=int, default=4096, help="polynomials per GPU (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--only", default="", help="comma list of sections to run besides the headline: "
                    "fft,single,t64,t200,fft20,mfft,encode (default: all)")
    ap.add_argument("--skip-configs", action="store_true", help="headline + fft only")
    ap.add_argument("--small", action="store_true", help="shrink the side configurations (plumbing check, not a measurement)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_counters(kernel, tag=None):
    """per-unit counters of `kernel` from the committed ncu capture (profiles/traffic.json, made by
    tools/make_traffic.py from an `ncu --set full` export): DRAM bytes and alu-pipe warp instructions per
    codeword / field point; None when the kernel has no capture.  A side configuration (`tag`, e.g. "t64")
    only uses counters captured on that configuration ("kernel@tag"): per-unit work depends on the code."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        return d.get(f"{kernel}@{tag}") if tag else d.get(kernel)
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """samples SM clock / throttle reasons with NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# inputs (same generator and seeds for both arms)
# ---------------------------------------------------------------------------------------------
def make_decode_input(encode, data_bits, length, t, batch, seed):
    """pool of 64 encoded codewords -> `batch` received words (XOR of 4 pool entries + errors)"""
    bits = (synth.splitmix64(seed, 64 * data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([encode(b) for b in bits.reshape(64, -1)])
    return synth.corrupt_codewords(pool, length, t, batch, seed, beyond_fraction=0.01)


def decode_workload_name(length, t, batch, m):
    return (f"batched BCH decode GF(2^{m}) n={length} t={t} (BASELINE configs[2]); {batch} codewords per GPU")


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores (oracle/_ref only)
# ---------------------------------------------------------------------------------------------
def cpu_reference_decode(length, t, words, cores):
    """decode `words` with the unmodified reference (oracle/_ref) on `cores` threads; falls back to the
    single-threaded C restatement when oracle/_ref is not present.  Returns (codewords/s, kind, ok,
    corrected, decoded words, threads actually used)."""
    from oracle import pyoracle
    if pyoracle.ref_available():
        rc = pyoracle.Ref().code(length, t)
        ok, corr, out, dt = rc.decode_batch_mt(words, cores)
        return len(words) / dt, "reference", ok, corr, out, cores
    pc = pyoracle.Port().code(length, t)
    t0 = time.perf_counter()
    ok, corr, out = pc.decode_batch(words)
    return len(words) / (time.perf_counter() - t0), "port", ok, corr, out, 1


def cpu_reference_fft(m, polys, cores, mult=False):
    from oracle import pyoracle
    n = 1 << m
    if pyoracle.ref_available():
        R = pyoracle.Ref()
        out, dt = (R.multiplicative_fft_batch_mt(m, polys, cores) if mult
                   else R.additive_fft_batch_mt(m, polys, n - 1, cores))
        return len(polys) * n / dt, "reference", out, cores
    P = pyoracle.Port()
    t0 = time.perf_counter()
    out = np.stack([(P.mult_fft(m, p[: n - 1]) if mult else P.additive_fft(m, p, n - 1)[0]) for p in polys])
    return len(polys) * n / (time.perf_counter() - t0), "port", out, 1


def reference_code(length, t):
    """encoder for building inputs on the reference arm: the unmodified reference (or the restatement
    when oracle/_ref is absent) -- none of this repository's libraries are loaded on that arm"""
    from oracle import pyoracle
    return (pyoracle.Ref() if pyoracle.ref_available() else pyoracle.Port()).code(length, t)


def run_reference(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    sample = min(args.batch, 12000 * cores)
    enc = reference_code(args.length, args.t)
    words, _ = make_decode_input(enc.encode, args.length - enc.gen_degree, args.length, args.t, sample, 0xC3)
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        rate, kind, *_ = cpu_reference_decode(args.length, args.t, words, cores)
        if i >= args.warmup:
            times.append(sample / rate)
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "codewords/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_dict(args, sample_note=f"each step decodes a {sample}-codeword sample of the workload"),
        "cpu_baseline": {"value": value, "unit": "codewords/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} codewords per step, bch_decode on {cores} threads"},
        "e2e": {"value": value, "unit": "codewords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def config_dict(args, sample_note=None):
    m = max(2, int(args.length).bit_length())        # bch_gen_code: smallest field with 2^m >= length
    c = {"workload": decode_workload_name(args.length, args.t, args.batch, m),
         "codeword_bits": args.length, "t": args.t, "batch_per_gpu": args.batch,
         "errors": "0..t uniform per word, 1% of words t+1..t+3 (uncorrectable path)",
         "l2": "inputs (1 KiB per codeword) exceed the 126 MB L2; the work buffer is restored "
               "from a master copy between steps, outside the per-step event pair"}
    if sample_note:
        c["sample"] = sample_note
    return c


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """what every section needs: device, libraries, rank plumbing, roofline denominators"""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from bch_fft_b200 import _lib, api
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.dist, self._lib, self.api = torch, dist, _lib, api
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.lib = _lib.lib()
        self.host = api.Host()
        os.environ["BCHFFT_DEVICE"] = str(local_rank)       # the host library's device (e2e legs)
        self.peak_gbs, self.sm_mhz, self.peak_src = peaks()
        self.stream = torch.cuda.Stream(device=self.dev)
        self.cores = host_cores()
        self.cpu_legs = rank == 0 and world == 1 and not args.no_cpu_baseline
        self._pinned = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        tns = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tns, op=op)
        return float(tns.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def pinned(self, key, nbytes):
        """page-locked host buffers are expensive to create: one per role, grown on demand"""
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < nbytes:
            self._pinned[key] = None
            buf = self._pinned[key] = self.torch.empty(nbytes, dtype=self.torch.uint8).pin_memory()
        return buf[:nbytes]

    def time_steps(self, step, steps, warmup, clocks=None):
        """W untimed steps, then K steps each bracketed by CUDA events on the launching stream; barrier +
        synchronize on both sides; returns the summed device time in ms, max over ranks"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                step(False)
            self.barrier()
            evs = []

            def timed():
                for _ in range(steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    step(True, a, b)
                    evs.append((a, b))
                torch.cuda.synchronize()
            if clocks is not None:
                with clocks:
                    timed()
            else:
                timed()
            self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    def roofline(self, prof, units, bytes_per_unit, step_ms, unit_name, tag=None):
        """roofline of the dominant kernel: algorithmic bytes of one step / the kernel's device time per
        step (all of its launches), against the measured HBM peak; plus the integer-pipe view: alu-pipe
        warp instructions per unit (ncu capture) x units/s against 148 SMs x 64 lanes x SM clock"""
        ksum = sum(ms for ms, _ in prof.values()) or 1.0
        top = max(prof, key=lambda k: prof[k][0])
        top_ms, top_n = prof[top]
        achieved = units * bytes_per_unit / (top_ms * 1e-3) / 1e9
        cnt = ncu_counters(top, tag)
        r = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": self.peak_gbs, "unit": "GB/s",
             "frac": achieved / self.peak_gbs,
             "traffic": (cnt["dram_bytes_per_unit"] * units / top_n) if cnt else None,
             "peak_source": self.peak_src, f"algorithmic_bytes_per_{unit_name}": bytes_per_unit,
             f"{unit_name}s_per_step": units, "kernel_ms_per_step": top_ms, "launches_per_step": top_n,
             "launch_ms": top_ms / top_n,
             "kernel_share_of_step": {k: round(v[0] / ksum, 4) for k, v in prof.items()},
             "whole_step_frac": (units * bytes_per_unit / (step_ms * 1e-3) / 1e9) / self.peak_gbs}
        if cnt and cnt.get("alu_warp_inst_per_unit"):
            peak = SM_COUNT * ALU_LANES_PER_SM_CLK * self.sm_mhz * 1e6 / 1e12
            ach = cnt["alu_warp_inst_per_unit"] * 32 * units / (top_ms * 1e-3) / 1e12
            r["int_pipe"] = {"bound": "alu pipe (LOP3 / integer)", "achieved": ach, "peak": peak,
                             "unit": "T lane-op/s", "frac": ach / peak,
                             "alu_warp_inst_per_unit": cnt["alu_warp_inst_per_unit"],
                             "source": "profiles/traffic.json (ncu: sm__inst_executed_pipe_alu x cycles)"}
        return r


def decode_section(cx, length, t, B, steps, warmup, seed, headline=False, e2e_steps=3):
    """device-resident rate, roofline, e2e through bch_decode_batch, CPU reference + parity for one code"""
    torch, args = cx.torch, cx.args
    code = cx.host.bch_gen_code(length, 2 * t + 1)
    W = code.words
    words, _ = make_decode_input(code.encode, code.data_bits, length, t, B, seed + cx.rank)
    h_master = cx.pinned("dec_in", B * W * 8).view(torch.int64).view(B, W)
    h_master.copy_(torch.from_numpy(words.view(np.int64)))
    field = cx.host.create_binary_finite_field(code.m)
    dfield = cx.api.DeviceField(field, cx.local_rank)
    dcode = cx.api.DeviceCode(dfield, code)
    clocks = ClockSampler(cx.local_rank) if headline else None
    bytes_per_cw = 2 * W * 8 + 8           # codeword in + codeword out + 8 B status
    with torch.cuda.stream(cx.stream):
        d_master = h_master.to(cx.dev, non_blocking=True)
        d_work = torch.empty_like(d_master)
        d_ok = torch.empty(B, dtype=torch.uint8, device=cx.dev)
        d_corr = torch.empty(B, dtype=torch.int32, device=cx.dev)
    sp = cx.stream.cuda_stream

    def run():
        dcode.decode_dev(d_work.data_ptr(), B, d_ok.data_ptr(), d_corr.data_ptr(), sp)

    def step(timed, a=None, b=None):
        d_work.copy_(d_master)              # restore the received words (decode is in place): untimed
        if timed:
            a.record(cx.stream)
        run()
        if timed:
            b.record(cx.stream)

    cx.barrier()
    cx.lib.bchfft_launch_count_reset()
    total_ms = cx.time_steps(step, steps, warmup, clocks)
    launches = cx.lib.bchfft_launch_count()
    step_ms = total_ms / steps
    value = cx.world * B / (step_ms * 1e-3)
    with torch.cuda.stream(cx.stream):
        d_work.copy_(d_master)
        torch.cuda.synchronize()
        prof = cx._lib.profile(cx.lib, run)
        ok_gpu = d_ok.cpu().numpy()
        corr_gpu = d_corr.cpu().numpy().view(np.uint32)
    # ncu counters per codeword depend on the code: the headline uses the headline capture, t = 64 / 200 their own
    roofline = cx.roofline(prof, B, bytes_per_cw, step_ms, "codeword",
                           tag=None if (headline or (length, t) == (8191, 16)) else f"t{t}")
    roofline["note"] = ("integer/LOP3-bound bitsliced GF(2^m) arithmetic: the HBM fraction is low by construction; "
                        "int_pipe and profiles/*_ncu_summary_decode.txt give the pipe utilisation")
    out_head = None
    e2e = None
    if not args.no_e2e:
        ulp = C.POINTER(C.c_ulong)
        h_work = cx.pinned("dec_work", B * W * 8).view(torch.int64).view(B, W)
        h_ok = cx.pinned("dec_ok", B)
        h_corr = cx.pinned("dec_corr", B * 4).view(torch.int32)
        tsum = 0.0
        for i in range(1 + e2e_steps):
            h_work.copy_(h_master)
            cx.barrier()
            t0 = time.perf_counter()
            rc = cx.host.lib.bch_decode_batch(C.byref(code.c), C.cast(h_work.data_ptr(), ulp), B,
                                              C.cast(h_ok.data_ptr(), C.POINTER(C.c_ubyte)),
                                              C.cast(h_corr.data_ptr(), C.POINTER(C.c_uint32)))
            dt = time.perf_counter() - t0
            assert rc == 0, cx.host.lib.bchfft_last_error()
            if i:
                tsum += cx.max_over_ranks(dt)
        assert (h_ok.numpy() == ok_gpu).all()
        # what bch_decode_batch brings back (cabi_bch.cu, BCHFFT_D2H_MODE): verdicts + error positions from
        # 8192 codewords on (host threads flip the bits in the caller's array), whole words below that
        mode = os.environ.get("BCHFFT_D2H_MODE") or ("positions" if B >= 8192 else "words")
        d2h = B * (1 + 4 + 4 * (t + 1)) if mode == "positions" else B * (W * 8 + 1 + 4)
        e2e = {"value": cx.world * B * e2e_steps / tsum, "unit": "codewords/s",
               "h2d_bytes_per_step": B * W * 8, "d2h_bytes_per_step": d2h, "result_mode": mode,
               "steps": e2e_steps, "api": "bch_decode_batch (host C library, include/bch.h)",
               "timer": "host wall clock around the synchronous call, max over ranks"}
        out_head = h_work.numpy().view(np.uint64)
    cpu = parity = None
    if cx.cpu_legs:
        per_core = {16: 12000, 64: 1500, 200: 400}.get(t, max(100, 200000 // (t * t)))
        sample = min(B, per_core * cx.cores)
        rate, kind, rok, rcorr, rout, used = cpu_reference_decode(length, t, words[:sample], cx.cores)
        cpu = {"value": rate, "unit": "codewords/s", "cores": used, "kind": kind,
               "sample": f"first {sample} codewords of the same batch, bch_decode on {used} host threads"}
        parity = {"codewords": sample,
                  "ok_mismatches": int((rok != ok_gpu[:sample]).sum()),
                  "corrected_mismatches": int((rcorr != corr_gpu[:sample]).sum())}
        if out_head is not None:
            parity["word_mismatches"] = int((rout != out_head[:sample]).any(axis=1).sum())
    stats = {"ok_fraction": cx.sum_over_ranks(float(ok_gpu.mean())) / cx.world,
             "mean_corrected": cx.sum_over_ranks(float(corr_gpu.mean())) / cx.world}
    dcode.close()
    dfield.close()
    del d_master, d_work, d_ok, d_corr
    return {"metric": METRIC, "value": value, "unit": "codewords/s", "ms_per_step": step_ms, "steps": steps,
            "config": {"workload": decode_workload_name(length, t, B, code.m), "codeword_bits": length, "t": t,
                       "batch_per_gpu": B},
            "gpu_launches": int(launches), "clocks": clocks.summary() if clocks else None,
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity": parity, "decode_stats": stats}


def fft_section(cx, m, Bf, steps, warmup, seed, mult=False, e2e_steps=2):
    """batched additive (or multiplicative) FFT: device-resident rate, roofline, e2e through the host
    library's batch entry point, CPU reference + parity"""
    torch, args = cx.torch, cx.args
    n = 1 << m
    field = cx.host.create_binary_finite_field(m)
    dfield = cx.api.DeviceField(field, cx.local_rank)
    polys = synth.random_polys(m, Bf, seed + cx.rank)
    h_in = cx.pinned("fft_in", Bf * n * 4).view(torch.int32).view(Bf, n)
    h_in.copy_(torch.from_numpy(polys.view(np.int32)))
    with torch.cuda.stream(cx.stream):
        d_in = h_in.to(cx.dev, non_blocking=True)
        d_out = torch.empty_like(d_in)
    sp = cx.stream.cuda_stream

    def run():
        if mult:
            dfield.multiplicative_fft_dev(d_in.data_ptr(), d_out.data_ptr(), n - 1, Bf, sp)
        else:
            dfield.additive_fft_dev(d_in.data_ptr(), d_out.data_ptr(), n - 1, Bf, sp)

    def step(timed, a=None, b=None):
        if timed:
            a.record(cx.stream)
        run()
        if timed:
            b.record(cx.stream)

    cx.lib.bchfft_launch_count_reset()
    total_ms = cx.time_steps(step, steps, warmup)
    launches = cx.lib.bchfft_launch_count()
    step_ms = total_ms / steps
    with torch.cuda.stream(cx.stream):
        prof = cx._lib.profile(cx.lib, run)
        res_gpu = d_out.cpu().numpy().view(np.uint32)
    value = cx.world * Bf * n / (step_ms * 1e-3)
    bytes_per_point = 4 if m <= 16 else 8      # 2 x element width
    roofline = cx.roofline(prof, Bf * n, bytes_per_point, step_ms, "point")
    kind = "multiplicative" if mult else "additive"
    e2e = None
    if not args.no_e2e:
        u32p = C.POINTER(C.c_uint32)
        h_out = cx.pinned("fft_out", Bf * n * 4).view(torch.int32).view(Bf, n)
        fn = cx.host.lib.multiplicative_fft_batch if mult else cx.host.lib.additive_fft_batch
        tsum = 0.0
        for i in range(1 + e2e_steps):
            cx.barrier()
            t0 = time.perf_counter()
            rc = fn(C.byref(field.c), C.cast(h_in.data_ptr(), u32p), C.cast(h_out.data_ptr(), u32p), n - 1, Bf)
            dt = time.perf_counter() - t0
            assert rc == 0, cx.host.lib.bchfft_last_error()
            if i:
                tsum += cx.max_over_ranks(dt)
        assert (h_out.numpy().view(np.uint32)[:, : n - 1] == res_gpu[:, : n - 1]).all()
        e2e = {"value": cx.world * Bf * n * e2e_steps / tsum, "unit": "points/s",
               "h2d_bytes_per_step": Bf * n * 4, "d2h_bytes_per_step": Bf * (n - 1) * 4, "steps": e2e_steps,
               "api": f"{kind}_fft_batch (host C library, include/{kind}_fft.h)",
               "timer": "host wall clock around the synchronous call, max over ranks"}
    cpu = parity = None
    if cx.cpu_legs:
        sample = min(Bf, (24 * cx.cores if m <= 16 else cx.cores) if not mult else 4 * cx.cores)
        rate, ckind, rout, used = cpu_reference_fft(m, polys[:sample], cx.cores, mult)
        cpu = {"value": rate, "unit": "points/s", "cores": used, "kind": ckind,
               "sample": f"first {sample} polynomials of the same batch on {used} host threads"}
        parity = {"polynomials": sample,
                  "mismatching_points": int((rout[:, : n - 1] != res_gpu[:sample, : n - 1]).sum())}
    dfield.close()
    del d_in, d_out
    cfg_name = {(False, 16): "BASELINE configs[0] batched, SURVEY 8d C1", (False, 20): "BASELINE configs[3], SURVEY 8d C4",
                (True, 16): "BASELINE configs[4], SURVEY 8d C5"}.get((mult, m), "")
    return {"metric": f"{kind}_fft_field_points_per_s", "value": value, "unit": "points/s", "ms_per_step": step_ms,
            "steps": steps, "gpu_launches": int(launches),
            "config": {"workload": f"batched {kind} FFT GF(2^{m}), {Bf} random full-degree polynomials per GPU ({cfg_name})",
                       "m": m, "batch_per_gpu": Bf, "io": "u32 coefficients in, u32 evaluations out (reference ABI)",
                       "l2": f"inputs and outputs (2 x {Bf * n * 4 >> 20} MiB) exceed the L2"},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity": parity}


def encode_section(cx, length, t, B, steps, warmup):
    """bch_encode_batch device-resident, against the reference's bch_encode on all cores"""
    torch = cx.torch
    code = cx.host.bch_gen_code(length, 2 * t + 1)
    k, W = code.data_bits, code.words
    dw = (k + 63) // 64
    data = synth.splitmix64(0xE0 + cx.rank, B * dw).reshape(B, dw)
    field = cx.host.create_binary_finite_field(code.m)
    dfield = cx.api.DeviceField(field, cx.local_rank)
    dcode = cx.api.DeviceCode(dfield, code)
    with torch.cuda.stream(cx.stream):
        d_data = torch.from_numpy(data.view(np.int64)).to(cx.dev)
        d_cw = torch.empty((B, W), dtype=torch.int64, device=cx.dev)
    sp = cx.stream.cuda_stream

    def step(timed, a=None, b=None):
        if timed:
            a.record(cx.stream)
        dcode.encode_dev(d_data.data_ptr(), dw, k, d_cw.data_ptr(), B, sp)
        if timed:
            b.record(cx.stream)

    total_ms = cx.time_steps(step, steps, warmup)
    step_ms = total_ms / steps
    with torch.cuda.stream(cx.stream):
        torch.cuda.synchronize()
        cw_gpu = d_cw.cpu().numpy().view(np.uint64)
    bytes_per_cw = dw * 8 + W * 8
    achieved = B * bytes_per_cw / (step_ms * 1e-3) / 1e9
    cpu = parity = None
    if cx.cpu_legs:
        from oracle import pyoracle
        if pyoracle.ref_available():
            sample = min(B, 4000 * cx.cores)
            rcw, dt = pyoracle.Ref().code(length, t).encode_batch_mt(data[:sample], k, cx.cores)
            cpu = {"value": sample / dt, "unit": "codewords/s", "cores": cx.cores, "kind": "reference",
                   "sample": f"first {sample} data words, bch_encode on {cx.cores} host threads"}
            parity = {"codewords": sample, "word_mismatches": int((rcw != cw_gpu[:sample]).any(axis=1).sum())}
    dcode.close()
    dfield.close()
    return {"metric": "bch_codewords_encoded_per_s", "value": cx.world * B / (step_ms * 1e-3), "unit": "codewords/s",
            "ms_per_step": step_ms, "steps": steps,
            "config": {"workload": f"batched systematic encode n={length} t={t}, {B} data words per GPU",
                       "batch_per_gpu": B},
            "roofline": {"bound": "hbm", "kernel": "k_encode", "achieved": achieved, "peak": cx.peak_gbs, "unit": "GB/s",
                         "frac": achieved / cx.peak_gbs, "traffic": None,
                         "algorithmic_bytes_per_codeword": bytes_per_cw},
            "cpu_baseline": cpu, "parity": parity}


def single_call_section(cx):
    """BASELINE configs[0] / configs[1]: latency of ONE call of the reference's
    single-item entry points with host buffers -- evaluate_polynomial_additive_FFT over GF(2^16) and
    bch_decode at n = 65535 with t errors -- through the drop-in library (GPU, copies included; this IS the
    end-to-end number) and through the unmodified reference (oracle/_ref, one thread: the call is
    single-threaded by definition), outputs compared; kernel breakdown and roofline from one profiled call."""
    from oracle import pyoracle
    host = cx.host
    ref = pyoracle.Ref() if (cx.cpu_legs and pyoracle.ref_available()) else None
    u32p, ulp = C.POINTER(C.c_uint32), C.POINTER(C.c_ulong)
    res = {}
    reps = 20
    m, n = 16, 1 << 16
    f = host.create_binary_finite_field(m)
    poly = synth.random_polys(m, 1, 0xC1)[0]
    want, cpu_ms = None, None
    if ref is not None:
        ts = []
        for _ in range(3):
            want, dt = ref.additive_fft_batch_mt(m, poly[None, :], n - 1, 1)
            ts.append(dt)
        cpu_ms = 1e3 * min(ts)
    out = np.zeros(n, np.uint32)

    def fft_call():
        work = poly.copy()
        t0 = time.perf_counter()
        host.lib.evaluate_polynomial_additive_FFT(C.byref(f.c), work.ctypes.data_as(u32p), out.ctypes.data_as(u32p), n - 1)
        return time.perf_counter() - t0

    ts = sorted([fft_call() for _ in range(3 + reps)][3:])
    prof = cx._lib.profile(cx.lib, fft_call)
    med = ts[len(ts) // 2]
    res["additive_fft_gf2_16"] = {
        "metric": "single_call_latency_ms", "value": 1e3 * med, "unit": "ms", "higher_is_better": False,
        "gpu_ms_median": 1e3 * med, "gpu_ms_min": 1e3 * ts[0],
        "api": "evaluate_polynomial_additive_FFT (host buffers, copies included = end to end)",
        "e2e": {"value": n / med, "unit": "points/s", "h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": 2 * n * 4},
        "roofline": cx.roofline(prof, n, 4, 1e3 * med, "point") if prof else None,
        "cpu_baseline": None if cpu_ms is None else {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "reference",
                                                     "sample": "the same polynomial, best of 3"},
        "parity": None if want is None else {"points": n - 1, "mismatching_points": int((out[: n - 1] != want[0]).sum())}}
    for t in (100, 1000):
        length = 65535
        code = host.bch_gen_code(length, 2 * t + 1)
        word = code.encode((synth.splitmix64(0xC2 + t, code.data_bits) & np.uint64(1)).astype(np.uint32))
        rng = np.random.default_rng(0xC2 + t)
        for p in rng.choice(np.arange(1, length), t, replace=False):
            word[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
        cpu_ms, ref_out = None, None
        if ref is not None:
            rc = ref.code(length, t)
            ts = []
            for _ in range(3):
                ok_r, corr_r, w_r, dt = rc.decode_batch_mt(word[None, :], 1)
                ts.append(dt)
            cpu_ms = 1e3 * min(ts)
            ref_out = (int(ok_r[0]), int(corr_r[0]), w_r[0])
        corr = C.c_uint32(0)
        state = {}

        def dec_call():
            work = word.copy()
            t0 = time.perf_counter()
            state["ok"] = host.lib.bch_decode(C.byref(code.c), work.ctypes.data_as(ulp), None, C.byref(corr))
            dt = time.perf_counter() - t0
            state["work"] = work
            return dt

        ts = sorted([dec_call() for _ in range(3 + reps)][3:])
        prof = cx._lib.profile(cx.lib, dec_call)
        med = ts[len(ts) // 2]
        ok, work = state["ok"], state["work"]
        wbytes = code.words * 8
        res[f"bch_decode_n65535_t{t}"] = {
            "metric": "single_call_latency_ms", "value": 1e3 * med, "unit": "ms", "higher_is_better": False,
            "gpu_ms_median": 1e3 * med, "gpu_ms_min": 1e3 * ts[0],
            "api": "bch_decode (host buffers, copies included = end to end)",
            "e2e": {"value": 1.0 / med, "unit": "codewords/s", "h2d_bytes_per_step": wbytes, "d2h_bytes_per_step": wbytes + 5},
            "roofline": cx.roofline(prof, 1, 2 * wbytes + 8, 1e3 * med, "codeword") if prof else None,
            "cpu_baseline": None if cpu_ms is None else {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "reference",
                                                         "sample": "the same word, best of 3"},
            "ok": int(ok), "corrected": int(corr.value),
            "parity": None if ref_out is None else {
                "codewords": 1, "word_mismatches": int(not (ref_out[0] == int(ok) and ref_out[1] == int(corr.value)
                                                            and (ref_out[2] == work).all()))}}
    return res


def run_ours(args, rank, world, local_rank):
    cx = Ctx(args, rank, world, local_rank)
    want = set(filter(None, args.only.split(","))) if args.only else {"fft", "single", "t64", "t200", "fft20", "mfft", "encode"}
    if args.skip_configs:
        want &= {"fft"}
    head = decode_section(cx, args.length, args.t, args.batch, args.steps, args.warmup, 0xC3, headline=True,
                          e2e_steps=max(1, min(args.steps, 3)))
    side_steps = max(2, min(args.steps, 5))
    sm = 32 if args.small else 1            # --small: plumbing check only
    fft = None
    if args.fft_batch > 0 and "fft" in want:
        fft = fft_section(cx, args.fft_m, args.fft_batch, args.steps, args.warmup, 0xC1)
    configs = {}
    if "single" in want and rank == 0:
        configs["single_call"] = single_call_section(cx)
    if world > 1:
        cx.barrier()
    if "t64" in want:
        configs["decode_n8191_t64"] = decode_section(cx, 8191, 64, 262144 // sm, side_steps, 3, 0xC64, e2e_steps=2)
    if "t200" in want:
        configs["decode_n8191_t200"] = decode_section(cx, 8191, 200, 131072 // sm, side_steps, 3, 0xC200, e2e_steps=2)
    if "fft20" in want:
        configs["additive_fft_gf2_20"] = fft_section(cx, 20, 256 // sm, side_steps, 3, 0xC4)
    if "mfft" in want:
        configs["multiplicative_fft_gf2_16"] = fft_section(cx, 16, 4096 // sm, side_steps, 3, 0xC5, mult=True)
    if "encode" in want:
        configs["encode_n8191_t16"] = encode_section(cx, 8191, 16, 1_000_000 // sm, side_steps, 3)
    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "codewords/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config_dict(args),
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
            "e2e": head["e2e"], "cpu_baseline": head["cpu_baseline"], "parity": head["parity"],
            "decode_stats": head["decode_stats"], "fft": fft, "configs": configs,
        }
        emit(line)


_RESULT_FD = None


def emit(line):
    """the one JSON line goes to the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse()
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner
    # under torchrun) are diverted to stderr for the whole run
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
