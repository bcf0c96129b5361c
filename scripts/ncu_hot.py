"""Print the hottest SASS instructions (by warp-stall samples) from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
hdr = rows[hdr_i]
si = hdr.index("Source")
j = next(k for k, h in enumerate(hdr) if "Warp Stall Sampling (All" in h)
data = []
for r in rows[hdr_i + 1:]:
    try:
        data.append((int(r[j]), r[si][:120]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for n, src in sorted(data, reverse=True)[:top]:
    print(f"{n:7d} {100 * n / tot:5.1f}%  {src}")
