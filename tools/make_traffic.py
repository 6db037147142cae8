"""profiles/traffic.json from an `ncu --page raw --csv` export: measured DRAM bytes (read + write)
per launch of every kernel, divided by the units (codewords / field points) one launch processed.
usage: make_traffic.py raw.csv decode_codewords_per_launch fft_points_per_launch "command" > traffic.json"""
import csv
import json
import sys

DECODE = ("k_syndrome", "k_syndrome_p", "k_syn_finish", "k_bm", "k_bucket", "k_elp_fwd", "k_elp_fwd_u",
          "k_locate", "k_locate_u", "k_correct")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ni].split("(")[0].replace("void ", "").replace("bchfft::", "").split("<")[0]
    b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
    acc.setdefault(name, []).append(b)
out = {}
for k, v in acc.items():
    dec = k in DECODE
    u = float(sys.argv[2]) if dec else float(sys.argv[3])
    out[k] = {"dram_bytes_per_launch_mean": sum(v) / len(v), "units_per_launch": u,
              "unit": "codeword" if dec else "point", "dram_bytes_per_unit": sum(v) / len(v) / u,
              "launches_sampled": len(v)}
out["_source"] = ("ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                  + sys.argv[1] + "; command: " + sys.argv[4])
json.dump(out, sys.stdout, indent=1)
