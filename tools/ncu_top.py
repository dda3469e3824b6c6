"""Top stalled SASS instructions per kernel from `ncu --page source --csv`. Usage: ncu_top.py file.csv [kernel_index] [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
s, e = ks[which], ks[which + 1]
print(rows[s][1][:60], "kernels:", len(ks) - 1)
H = rows[s + 1]
data = [r for r in rows[s + 2:e] if len(r) > 5]
isamp, isrc, iex = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for h in stalls:
        agg[h[6:]] = agg.get(h[6:], 0) + int(r[H.index(h)] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n]:
    st = {h[6:]: int(r[H.index(h)] or 0) for h in stalls}
    st = {k: v for k, v in st.items() if v > 0.15 * int(r[isamp] or 1)}
    print(r[isamp], r[iex], r[isrc][:86], st)
