"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel over the last step.
usage: python tools/launch_summary.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
last = L[-(len(L) // steps):]
agg = collections.OrderedDict()
for k, v in last:
    k = re.sub(r"\(.*", "", k)
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-64s %8.1f us %3d launches %5.1f%%" % (k[:64], v / 1000, c, 100 * v / tot))
print("launches %d, total %.1f us" % (len(last), tot / 1000))
