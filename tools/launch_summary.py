"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
ui = hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
acc = {}
order = []
for r in rows[hi + 1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("bchfft::", "")
    us = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)
    if name not in acc:
        acc[name] = [0, 0.0, r[gi], r[bi]]
        order.append(name)
    acc[name][0] += 1
    acc[name][1] += us
DEC = ("k_syndrome", "k_syn_finish", "k_bm", "k_bucket", "k_elp_fwd", "k_locate", "k_correct")
for title, sel in (("decode step", lambda n: n.startswith(DEC)), ("additive FFT step", lambda n: not n.startswith(DEC))):
    names = [n for n in order if sel(n)]
    tot = sum(acc[n][1] for n in names) or 1.0
    print(title)
    for n in names:
        c, us, g, b = acc[n]
        print(f"  {n:22s} launches={c:3d} total={us:9.1f} us avg={us / c:8.1f} us share={us / tot:.3f} grid={g} block={b}")
