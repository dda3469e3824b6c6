"""Stall samples per SASS region of the GEMM kernel, in address order, with marker instructions (UTMALDG / UTCHMMA / LDTM /
UTMASTG / SYNCS) so that each mbarrier spin loop can be attributed to its role.  Usage: ncu_roles.py src.csv kernel_index [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
which = int(sys.argv[2]); minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
s, e = ks[which], ks[which + 1]
H = rows[s + 1]
data = [r for r in rows[s + 2:e] if len(r) > 5]
isamp, isrc, iex = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print(rows[s][1][:70], "samples", tot)
MARK = ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "UTMASTG", "SYNCS", "ELECT", "RED.", "ATOM", "BAR.", "UCGABAR", "EXIT", "STS", "LDS", "STG", "LDG")
last_mark = None
for idx, r in enumerate(data):
    n = int(r[isamp] or 0)
    src = r[isrc].strip()
    is_mark = any(m in src for m in MARK)
    if n >= minpct * tot / 100.0:
        st = {h[6:]: int(r[H.index(h)] or 0) for h in stalls}
        st = {k: v for k, v in st.items() if v > 0.15 * max(n, 1)}
        print("%5d %5.1f%% ex=%-8s #%-5d %s %s" % (n, 100.0 * n / tot, r[iex], idx, src[:80], st))
    elif is_mark and src.split()[0 if not src.startswith('@') else 1][:7] != last_mark:
        print("              ex=%-8s #%-5d   . %s" % (r[iex], idx, src[:80]))
        last_mark = src.split()[0 if not src.startswith('@') else 1][:7]
