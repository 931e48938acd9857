"""Summarise an `ncu --page source --csv` dump: top SASS lines by stall samples, with reasons."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col["# Samples"]])
    except ValueError:
        continue
    data.append((n, r))
tot = sum(n for n, _ in data)
print(f"total samples {tot}, instructions {len(data)}")
agg = {h: 0 for h in stall_cols}
for n, r in data:
    for h in stall_cols:
        try:
            agg[h] += int(r[col[h]])
        except ValueError:
            pass
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
for i, (n, r) in enumerate(data):
    r.append(i)
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    reasons = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100*n/tot:5.1f}%  #{r[-1]:5d} exec={r[col['Instructions Executed']]:>8s}  {r[col['Source']].strip()[:70]:70s} {reasons}")
