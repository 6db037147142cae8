"""profiles/traffic.json from `ncu --page raw --csv` exports (one row per captured launch): per kernel
the measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the alu-pipe warp instructions
(sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active x sm__cycles_active.avg x 148 SMs x 2 warp
instructions per clock: the alu pipe issues one warp instruction every two cycles per SMSP, four SMSPs) per
launch, divided by the units (codewords / field points) one launch processed.  bench.py reads the file for
`roofline.traffic` and `roofline.int_pipe`.
usage: make_traffic.py out.json "command" raw1.csv:units_per_launch[:kernel_regex[:tag]] [raw2.csv:units ...]
The regex is searched in the kernel name with its template arguments ("k_bm<13, 1, 8>"); with a tag the entry
is stored as "kernel@tag" (a side configuration's own counters: bench.py looks "kernel@tag" up first).
A later (file, kernel) entry replaces an earlier one; existing entries of out.json for kernels not captured
this time are kept (their _source line stays in "_history")."""
import csv
import json
import os
import re
import sys

DECODE = ("k_syndrome", "k_syndrome_p", "k_syn_finish", "k_bm", "k_bm_wide", "k_bucket", "k_elp_fwd", "k_elp_fwd_u",
          "k_locate", "k_locate_u", "k_correct", "k_encode", "k_encode_wide", "k_unpack_bits", "k_syn_from_fft")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(x):
    return float(x.replace(",", ""))


def main():
    out_path, command = sys.argv[1], sys.argv[2]
    out = {}
    if os.path.exists(out_path):
        out = json.load(open(out_path))
    hist = out.get("_history", [])
    if out.get("_source"):
        hist.append(out["_source"])
    for spec in sys.argv[3:]:
        parts = spec.split(":")
        path, units = parts[0], float(parts[1])
        rx = re.compile(parts[2]) if len(parts) > 2 and parts[2] else None
        tag = parts[3] if len(parts) > 3 else ""
        rows = list(csv.reader(open(path)))
        hdr, unit_row = rows[0], rows[1]
        ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        ai = hdr.index("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
        ci = hdr.index("sm__cycles_active.avg")
        ti = hdr.index("gpu__time_duration.sum")
        acc = {}
        for r in rows[2:]:
            full = r[ni].split("(")[0].replace("void ", "").replace("bchfft::", "")
            name = full.split("<")[0]
            if rx and not rx.search(full):
                continue
            if tag:
                name = name + "@" + tag
            b = num(r[ri]) * SCALE[unit_row[ri]] + num(r[wi]) * SCALE[unit_row[wi]]
            alu = num(r[ai]) / 100.0 * num(r[ci]) * 148 * 2.0
            acc.setdefault(name, []).append((b, alu, num(r[ai]), num(r[ti])))
        for k, v in acc.items():
            n = len(v)
            out[k] = {"dram_bytes_per_launch_mean": sum(x[0] for x in v) / n, "units_per_launch": units,
                      "unit": "codeword" if k.split("@")[0] in DECODE else "point",
                      "dram_bytes_per_unit": sum(x[0] for x in v) / n / units,
                      "alu_warp_inst_per_unit": sum(x[1] for x in v) / n / units,
                      "alu_pipe_pct_of_peak_active": sum(x[2] for x in v) / n,
                      "duration_ns_under_ncu": sum(x[3] for x in v) / n,
                      "launches_sampled": n, "capture": os.path.basename(path)}
    out["_source"] = ("ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum and "
                      "alu-pipe instructions per launch; command: " + command + "; files: " + " ".join(sys.argv[3:]))
    out["_history"] = hist
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
