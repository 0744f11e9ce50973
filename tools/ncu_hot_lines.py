"""Top stall-sample instructions of one kernel from an ncu report's source page.
usage: ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; python tools/ncu_hot_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
si, src = H.index("Warp Stall Sampling (All Samples)"), H.index("Source")
data = []
for i, r in enumerate(rows[hi + 1:]):
    if len(r) > si and r[si].isdigit():
        data.append((int(r[si]), i, r[src]))
tot = sum(d[0] for d in data)
print(rows[0][1] if len(rows[0]) > 1 else "", "total samples", tot)
for n, i, sr in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%%  #%d %s" % (n, 100.0 * n / tot, i, sr.strip()[:100]))
