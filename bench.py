#!/usr/bin/env python
"""Benchmark of the hot path: batched BCH decode (codewords/s, the headline) and batched additive
FFT (field points/s, reported in the same line under "fft").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU code, all cores

A "step" is one pass of the hot path over one batch of synthetic input resident in HBM.  The
workload is BASELINE.json configs[2] (batched BCH decode over GF(2^13), n = 8191, 10^6 synthetic
codewords per GPU with injected errors) -- configs[1] is a single decode and carries no
throughput.  Multi-GPU: one process per GPU (torchrun), each rank decodes its own range of
codewords, no collective on the data path ("weak" scaling: per-GPU batch fixed).
"""
import argparse
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
CW_BYTES_8191 = 2056          # 1 KiB in + 1 KiB out + 8 B status per codeword


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--length", type=int, default=8191)
    ap.add_argument("--t", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1_000_000, help="codewords per GPU")
    ap.add_argument("--fft-m", type=int, default=16)
    ap.add_argument("--fft-batch", type=int, default=4096, help="polynomials per GPU (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extras", action="store_true",
                    help="also time BASELINE configs[3] (additive FFT GF(2^20), 256 polynomials) and "
                         "configs[4] (multiplicative FFT GF(2^16)) device-resident")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel, units_per_launch):
    """DRAM bytes per launch of `kernel` = bytes per unit measured by ncu (dram__bytes_read.sum +
    dram__bytes_write.sum, profiles/traffic.json, produced by tools/ncu_summary.py's input) x the
    units this launch processes; None when no capture exists for the kernel"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            ent = json.load(f).get(kernel)
        return ent["dram_bytes_per_unit"] * units_per_launch if ent else None
    except Exception:
        return None


def make_decode_input(code, length, t, batch, seed):
    bits = (synth.splitmix64(seed, 64 * code.data_bits) & np.uint64(1)).astype(np.uint32)
    pool = np.stack([code.encode(b) for b in bits.reshape(64, -1)])
    return synth.corrupt_codewords(pool, length, t, batch, seed, beyond_fraction=0.01)


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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_decode(length, t, words, cores):
    """decode `words` with the unmodified reference (oracle/_ref) on `cores` threads; falls back
    to the single-threaded C restatement when oracle/_ref is not present.  Returns
    (codewords/s, kind, ok, corrected, decoded words, threads actually used)."""
    from oracle import pyoracle
    if pyoracle.ref_available():
        rc = pyoracle.Ref().code(length, t)
        ok, corr, out, dt = rc.decode_batch_mt(words, cores)
        return len(words) / dt, "reference", ok, corr, out, cores
    pc = pyoracle.Port().code(length, t)
    t0 = time.perf_counter()
    ok, corr, out = pc.decode_batch(words)
    return len(words) / (time.perf_counter() - t0), "port", ok, corr, out, 1


def cpu_reference_fft(m, polys, cores):
    from oracle import pyoracle
    n = 1 << m
    if pyoracle.ref_available():
        out, dt = pyoracle.Ref().additive_fft_batch_mt(m, polys, n - 1, cores)
        return len(polys) * n / dt, "reference", out, cores
    P = pyoracle.Port()
    t0 = time.perf_counter()
    out = np.stack([P.additive_fft(m, p, n - 1)[0] for p in polys])
    return len(polys) * n / (time.perf_counter() - t0), "port", out, 1


def run_reference(args, rank):
    if rank != 0:
        return
    from bch_fft_b200 import api  # host library only for code construction / encoding of inputs
    cores = host_cores()
    sample = min(args.batch, 12000 * cores)
    # inputs come from the same generator / seed as our arm's rank 0
    try:
        code = api.Host().bch_gen_code(args.length, 2 * args.t + 1)
        enc = code
    except Exception:
        from oracle import pyoracle
        enc = pyoracle.Port().code(args.length, args.t)
        enc.data_bits = args.length - enc.gen_degree
    words, _ = make_decode_input(enc, args.length, args.t, sample, 0xC3)
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
    c = {"workload": f"batched BCH decode GF(2^13) n={args.length} t={args.t} "
                     f"(BASELINE configs[2]); {args.batch} codewords per GPU",
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
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from bch_fft_b200 import _lib, api

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.lib()
    host = api.Host()
    os.environ["BCHFFT_DEVICE"] = str(local_rank)       # the host library's device (e2e leg)
    peak_gbs, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.SUM)
        return float(tns.item())

    # ---- decode workload ------------------------------------------------------------------
    code = host.bch_gen_code(args.length, 2 * args.t + 1)
    B, W = args.batch, code.words
    words, nerr = make_decode_input(code, args.length, args.t, B, 0xC3 + rank)
    h_master = torch.from_numpy(words.view(np.int64)).pin_memory()
    field = host.create_binary_finite_field(code.m)
    dfield = api.DeviceField(field, local_rank)
    dcode = api.DeviceCode(dfield, code)
    stream = torch.cuda.Stream(device=dev)
    clocks = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        d_master = h_master.to(dev, non_blocking=True)
        d_work = torch.empty_like(d_master)
        d_ok = torch.empty(B, dtype=torch.uint8, device=dev)
        d_corr = torch.empty(B, dtype=torch.int32, device=dev)
        sp = stream.cuda_stream

        def step(timed):
            d_work.copy_(d_master)
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
            dcode.decode_dev(d_work.data_ptr(), B, d_ok.data_ptr(), d_corr.data_ptr(), sp)
            if timed:
                b.record(stream)
                return a, b

        for _ in range(args.warmup):
            step(False)
        barrier()
        lib.bchfft_launch_count_reset()
        with clocks:      # NVML polling thread runs for the whole timed region
            evs = [step(True) for _ in range(args.steps)]
            torch.cuda.synchronize()
        launches = lib.bchfft_launch_count()
        barrier()
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
        value = world * B * args.steps / (total_ms * 1e-3)

        # per-kernel breakdown of one more step (events around every launch)
        d_work.copy_(d_master)
        torch.cuda.synchronize()
        prof = _lib.profile(lib, lambda: dcode.decode_dev(d_work.data_ptr(), B, d_ok.data_ptr(),
                                                          d_corr.data_ptr(), sp))
        ok_gpu = d_ok.cpu().numpy()
        corr_gpu = d_corr.cpu().numpy().view(np.uint32)
        out_gpu_head = None

    ksum = sum(ms for ms, _ in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k][0])
    top_ms, top_n = prof[top]
    per_launch_units = B / top_n
    achieved = per_launch_units * CW_BYTES_8191 / (top_ms / top_n * 1e-3) / 1e9
    traffic = measured_traffic(top, per_launch_units)
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_codeword": CW_BYTES_8191,
                "codewords_per_launch": per_launch_units, "launch_ms": top_ms / top_n,
                "kernel_share_of_step": {k: round(v[0] / ksum, 4) for k, v in prof.items()},
                "whole_step_frac": (B * CW_BYTES_8191 / (total_ms / args.steps * 1e-3) / 1e9) / peak_gbs,
                "note": "integer/LOP3-bound bitsliced GF(2^13) arithmetic: the HBM fraction is low by "
                        "construction, see DESIGN.md section 4 and profiles/r1v_ncu_summary_decode.txt "
                        "for the ALU-pipe and shared-memory utilisation"}

    # ---- end to end through the drop-in host API (pinned host buffers, copies timed) ---------
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        h_work = torch.empty_like(h_master).pin_memory()
        h_ok = torch.empty(B, dtype=torch.uint8).pin_memory()
        h_corr = torch.empty(B, dtype=torch.int32).pin_memory()
        ulp = C.POINTER(C.c_ulong)
        e2e_steps = max(1, min(args.steps, 3))
        tsum = 0.0
        for i in range(1 + e2e_steps):
            h_work.copy_(h_master)
            barrier()
            t0 = time.perf_counter()
            rc = host.lib.bch_decode_batch(C.byref(code.c), C.cast(h_work.data_ptr(), ulp), B,
                                           C.cast(h_ok.data_ptr(), C.POINTER(C.c_ubyte)),
                                           C.cast(h_corr.data_ptr(), C.POINTER(C.c_uint32)))
            dt = time.perf_counter() - t0
            assert rc == 0, host.lib.bchfft_last_error()
            if i:
                tsum += max_over_ranks(dt)
        assert (h_ok.numpy() == ok_gpu).all()
        e2e = {"value": world * B * e2e_steps / tsum, "unit": "codewords/s",
               "h2d_bytes_per_step": B * W * 8, "d2h_bytes_per_step": B * (W * 8 + 1 + 4),
               "steps": e2e_steps, "api": "bch_decode_batch (host C library, include/bch.h)",
               "timer": "host wall clock around the synchronous call, max over ranks"}
        out_gpu_head = h_work.numpy().view(np.uint64)

    # ---- CPU baseline (rank 0, single GPU run only) + parity spot check ---------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        sample = min(B, 12000 * cores)
        rate, kind, rok, rcorr, rout, used = cpu_reference_decode(args.length, args.t, words[:sample], cores)
        cpu = {"value": rate, "unit": "codewords/s", "cores": used, "kind": kind,
               "sample": f"first {sample} codewords of the same batch, bch_decode on {used} host threads"}
        parity = {"codewords": sample,
                  "ok_mismatches": int((rok != ok_gpu[:sample]).sum()),
                  "corrected_mismatches": int((rcorr != corr_gpu[:sample]).sum())}
        if out_gpu_head is not None:
            parity["word_mismatches"] = int((rout != out_gpu_head[:sample]).any(axis=1).sum())

    stats = {"ok_fraction": sum_over_ranks(float(ok_gpu.mean())) / world,
             "mean_corrected": sum_over_ranks(float(corr_gpu.mean())) / world}

    # ---- additive FFT (second half of the metric) -------------------------------------------
    fft = None
    if args.fft_batch > 0:
        fft = run_fft(args, rank, world, local_rank, lib, host, peak_gbs, peak_src, barrier,
                      max_over_ranks, torch, api, _lib)

    extras = None
    if args.extras:
        extras = run_extras(args, rank, world, local_rank, lib, host, barrier, max_over_ranks, torch, api)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "codewords/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": config_dict(args), "gpu_launches": int(launches),
            "clocks": clocks.summary(), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "parity": parity, "decode_stats": stats, "fft": fft, "extras": extras,
        }
        emit(line)


def run_fft(args, rank, world, local_rank, lib, host, peak_gbs, peak_src, barrier, max_over_ranks,
            torch, api, _lib):
    m, Bf = args.fft_m, args.fft_batch
    n = 1 << m
    dev = torch.device("cuda", local_rank)
    field = host.create_binary_finite_field(m)
    dfield = api.DeviceField(field, local_rank)
    polys = synth.random_polys(m, Bf, 0xC1 + rank)
    h_in = torch.from_numpy(polys.view(np.int32)).pin_memory()
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        d_in = h_in.to(dev, non_blocking=True)
        d_out = torch.empty_like(d_in)
        sp = stream.cuda_stream

        def step():
            dfield.additive_fft_dev(d_in.data_ptr(), d_out.data_ptr(), n - 1, Bf, sp)

        for _ in range(args.warmup):
            step()
        barrier()
        evs = []
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        barrier()
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
        prof = _lib.profile(lib, step)
        res_gpu = d_out.cpu().numpy().view(np.uint32)
    value = world * Bf * n * args.steps / (total_ms * 1e-3)
    bytes_per_point = 4 if m <= 16 else 8      # 2 x element width
    ksum = sum(ms for ms, _ in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k][0])
    top_ms, top_n = prof[top]
    # every pass kernel is launched once per pass and per chunk of slabs; k_bitslice_in runs exactly
    # once per chunk, so one launch of the dominant kernel covers Bf * n / chunks points
    chunks = max(1, prof.get("k_bitslice_in", (0.0, 1))[1])
    points_per_launch = Bf * n / chunks
    achieved = points_per_launch * bytes_per_point / (top_ms / top_n * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": measured_traffic(top, points_per_launch),
                "peak_source": peak_src, "algorithmic_bytes_per_point": bytes_per_point, "points_per_launch": points_per_launch,
                "launch_ms": top_ms / top_n,
                "kernel_share_of_step": {k: round(v[0] / ksum, 4) for k, v in prof.items()},
                "whole_step_frac": (Bf * n * bytes_per_point / (total_ms / args.steps * 1e-3) / 1e9) / peak_gbs}
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        u32p = C.POINTER(C.c_uint32)
        h_out = torch.empty_like(h_in).pin_memory()
        tsum, steps = 0.0, max(1, min(args.steps, 3))
        for i in range(1 + steps):
            barrier()
            t0 = time.perf_counter()
            rc = host.lib.additive_fft_batch(C.byref(field.c), C.cast(h_in.data_ptr(), u32p),
                                             C.cast(h_out.data_ptr(), u32p), n - 1, Bf)
            dt = time.perf_counter() - t0
            assert rc == 0, host.lib.bchfft_last_error()
            if i:
                tsum += max_over_ranks(dt)
        assert (h_out.numpy().view(np.uint32)[:, : n - 1] == res_gpu[:, : n - 1]).all()
        e2e = {"value": world * Bf * n * steps / tsum, "unit": "points/s",
               "h2d_bytes_per_step": Bf * n * 4, "d2h_bytes_per_step": Bf * (n - 1) * 4,
               "steps": steps, "api": "additive_fft_batch (host C library, include/additive_fft.h)"}
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        sample = min(Bf, 24 * cores if m <= 16 else cores)
        rate, kind, rout, used = cpu_reference_fft(m, polys[:sample], cores)
        cpu = {"value": rate, "unit": "points/s", "cores": used, "kind": kind,
               "sample": f"first {sample} polynomials of the same batch on {used} host threads"}
        parity = {"polynomials": sample,
                  "mismatching_points": int((rout[:, : n - 1] != res_gpu[:sample, : n - 1]).sum())}
    return {"metric": "additive_fft_field_points_per_s", "value": value, "unit": "points/s",
            "ms_per_step": total_ms / args.steps,
            "config": {"workload": f"batched additive FFT GF(2^{m}), {Bf} random full-degree "
                                   f"polynomials per GPU (BASELINE configs[0] batched, SURVEY 8d C1)",
                       "m": m, "batch_per_gpu": Bf, "io": "u32 coefficients in, u32 evaluations out "
                       "(reference ABI)", "l2": "inputs and outputs (2 x 1 GiB) exceed the L2"},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity": parity}


def run_extras(args, rank, world, local_rank, lib, host, barrier, max_over_ranks, torch, api):
    """device-resident timings of the two remaining BASELINE configs (no e2e / CPU legs)"""
    dev = torch.device("cuda", local_rank)
    out = {}
    stream = torch.cuda.Stream(device=dev)
    for name, m, B, mult in (("additive_fft_gf2_20", 20, 256, False), ("multiplicative_fft_gf2_16", 16, 512, True)):
        n = 1 << m
        field = host.create_binary_finite_field(m)
        dfield = api.DeviceField(field, local_rank)
        polys = synth.random_polys(m, B, 0xC4 + rank)
        with torch.cuda.stream(stream):
            d_in = torch.from_numpy(polys.view(np.int32)).to(dev)
            d_out = torch.empty_like(d_in)
            sp = stream.cuda_stream

            def step():
                if mult:
                    dfield.multiplicative_fft_dev(d_in.data_ptr(), d_out.data_ptr(), n - 1, B, sp)
                else:
                    dfield.additive_fft_dev(d_in.data_ptr(), d_out.data_ptr(), n - 1, B, sp)

            for _ in range(2):
                step()
            barrier()
            steps = max(1, min(args.steps, 3))
            evs = []
            for _ in range(steps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                step()
                b.record(stream)
                evs.append((a, b))
            torch.cuda.synchronize()
            total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
        out[name] = {"value": world * B * n * steps / (total_ms * 1e-3), "unit": "points/s",
                     "ms_per_step": total_ms / steps, "m": m, "batch_per_gpu": B}
        dfield.close()
        del d_in, d_out
    if rank == 0 and world == 1:
        out["single_call"] = single_call_latency(host, args)
    return out


def single_call_latency(host, args):
    """BASELINE configs[0] / configs[1]: latency of ONE call of the reference's
    single-item entry points with host buffers -- evaluate_polynomial_additive_FFT over GF(2^16)
    and bch_decode at n = 65535 with t errors -- through the drop-in library (GPU, copies
    included) and through the unmodified reference (oracle/_ref, one thread), checked equal."""
    import ctypes as C
    from oracle import pyoracle
    ref = pyoracle.Ref() if pyoracle.ref_available() else None
    u32p, ulp = C.POINTER(C.c_uint32), C.POINTER(C.c_ulong)
    res = {}
    reps = 20
    # ---- one additive FFT, full degree, GF(2^16)
    m, n = 16, 1 << 16
    f = host.create_binary_finite_field(m)
    poly = synth.random_polys(m, 1, 0xC1)[0]
    want = None
    cpu_ms = None
    if ref is not None:
        ts = []
        for _ in range(3):
            want, dt = ref.additive_fft_batch_mt(m, poly[None, :], n - 1, 1)
            ts.append(dt)
        cpu_ms = 1e3 * min(ts)
    out = np.zeros(n, np.uint32)
    ts = []
    for _ in range(3 + reps):
        work = poly.copy()
        t0 = time.perf_counter()
        host.lib.evaluate_polynomial_additive_FFT(C.byref(f.c), work.ctypes.data_as(u32p),
                                                  out.ctypes.data_as(u32p), n - 1)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[3:])
    res["additive_fft_gf2_16"] = {"gpu_ms_median": 1e3 * ts[len(ts) // 2], "gpu_ms_min": 1e3 * ts[0],
                                  "reference_cpu_ms": cpu_ms,
                                  "equal": None if want is None else bool((out[: n - 1] == want[0]).all())}
    # ---- one decode at n = 65535, exactly t errors
    for t in (100, 1000):
        length = 65535
        code = host.bch_gen_code(length, 2 * t + 1)
        word = code.encode((synth.splitmix64(0xC2 + t, code.data_bits) & np.uint64(1)).astype(np.uint32))
        rng = np.random.default_rng(0xC2 + t)
        for p in rng.choice(np.arange(1, length), t, replace=False):
            word[p // 64] ^= np.uint64(1) << np.uint64(p % 64)
        cpu_ms = None
        ref_out = None
        if ref is not None:
            rc = ref.code(length, t)
            ts = []
            for _ in range(3):
                ok_r, corr_r, w_r, dt = rc.decode_batch_mt(word[None, :], 1)
                ts.append(dt)
            cpu_ms = 1e3 * min(ts)
            ref_out = (int(ok_r[0]), int(corr_r[0]), w_r[0])
        ts = []
        corr = C.c_uint32(0)
        for _ in range(3 + reps):
            work = word.copy()
            t0 = time.perf_counter()
            ok = host.lib.bch_decode(C.byref(code.c), work.ctypes.data_as(ulp), None, C.byref(corr))
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[3:])
        res[f"bch_decode_n65535_t{t}"] = {
            "gpu_ms_median": 1e3 * ts[len(ts) // 2], "gpu_ms_min": 1e3 * ts[0], "reference_cpu_ms": cpu_ms,
            "ok": int(ok), "corrected": int(corr.value),
            "equal": None if ref_out is None else bool(
                ref_out[0] == int(ok) and ref_out[1] == int(corr.value) and (ref_out[2] == work).all())}
    return res


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
