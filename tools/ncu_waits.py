"""Per kernel: stall samples inside each out-of-line mbarrier wait (attributed to its call sites' role), plus totals.
Usage: ncu_waits.py src.csv"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for which in range(len(ks) - 1):
    s, e = ks[which], ks[which + 1]
    pass
    H = rows[s + 1]; data = [r for r in rows[s + 2:e] if len(r) > 5]
    if not data: continue
    iaddr, isamp, isrc, iex = H.index("Address"), H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
    addr = [int(r[iaddr], 16) for r in data]; src = [r[isrc].strip() for r in data]; smp = [int(r[isamp] or 0) for r in data]
    idx = {a: i for i, a in enumerate(addr)}
    tot = sum(smp)
    def role(i):
        # nearest marker before/after
        for d in range(0, 4000):
            for j in (i - d, i + d):
                if 0 <= j < len(src):
                    t = src[j]
                    if "UTMALDG" in t: return "producer"
                    if "UTCHMMA" in t: return "mma"
                    if "LDTM" in t or "UTMASTG" in t: return "epilogue"
        return "?"
    calls = collections.defaultdict(list)
    for i, t in enumerate(src):
        m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)
        if m: calls[int(m.group(1), 16)].append(i)
    print("== kernel %d %s  samples %d" % (which, rows[s][1][22:66], tot))
    for tgt, sites in sorted(calls.items()):
        if tgt not in idx: continue
        i0 = idx[tgt]; n = 0; j = i0
        while j < len(src):
            n += smp[j]
            if src[j].startswith("RET") or " RET" in src[j]: break
            j += 1
        roles = collections.Counter(role(i) for i in sites)
        fast = sum(smp[i - k] for i in sites for k in range(0, 6) if i - k >= 0 and ("BRA" in src[i - k] or "TRYWAIT" in src[i - k]))
        print("   wait fn @%x: %5d samples (%4.1f%%) in slow path, %4d at call-site try; called from %s" % (tgt, n, 100.0 * n / tot, fast, dict(roles)))
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h[6:]: sum(int(r[H.index(h)] or 0) for r in data) for h in stalls}
    print("   stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.02 * tot})
