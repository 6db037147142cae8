"""Aggregate an `ncu --page source --csv` export: stall reasons over all samples and samples per opcode."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stalls}
ns = 0
byop = {}
for r in data:
    n = int(r[idx["# Samples"]] or 0)
    ns += n
    for h in stalls:
        tot[h] += int(r[idx[h]] or 0)
    f = r[idx["Source"]].split()
    op = " ".join(f[:2]) if f and f[0].startswith("@") else (f[0] if f else "")
    op = op.split(".")[0]
    d = byop.setdefault(op, [0, 0])
    d[0] += n
    d[1] += int(r[idx["Instructions Executed"]] or 0)
print("samples", ns)
for h, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    print(f"  {h:24s} {v:8d} {100 * v / max(ns, 1):5.1f}%")
ti = sum(v[1] for v in byop.values())
for op, (n, ie) in sorted(byop.items(), key=lambda x: -x[1][0])[:14]:
    print(f"  {op:14s} samples {n:7d} ({100 * n / max(ns, 1):4.1f}%)  warp-insts {ie:11d} ({100 * ie / max(ti, 1):4.1f}%)")
