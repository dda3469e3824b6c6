"""Summarise an `ncu --csv` launch list: one line per kernel launch with duration / tensor-pipe % / DRAM bytes."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
iid, ik, im, iv, iu = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
L = defaultdict(dict)
names = {}
for r in rows[h + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if u in ("ns", "nsecond"): v /= 1e3
    if u in ("msecond", "ms"): v *= 1e3
    if u == "Mbyte": v *= 1.0
    if u == "Kbyte": v /= 1e3
    if u == "Gbyte": v *= 1e3
    if u == "byte": v /= 1e6
    L[int(r[iid])][r[im]] = v
    names[int(r[iid])] = r[ik][:44]
tot = 0.0
for i in sorted(L):
    m = L[i]
    t = m.get("gpu__time_duration.sum", 0.0)
    tot += t
    print("%4d %-44s %9.1f us  tensor %5.1f%%  rd %7.1f MB  wr %7.1f MB" % (
        i, names[i], t, m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
        m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)))
print("total %.1f us" % tot)
