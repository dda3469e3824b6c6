"""Sum stall samples per role region (producer / epilogue / mma / tail) for every kernel in a source-page csv."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for which in range(len(ks) - 1):
    s, e = ks[which], ks[which + 1]
    H = rows[s + 1]
    data = [r for r in rows[s + 2:e] if len(r) > 5]
    if not data: continue
    isamp, isrc, iex = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    src = [r[isrc] for r in data]
    def first(tok, start=0):
        for i in range(start, len(src)):
            if tok in src[i]: return i
        return len(src)
    i_tma = first("UTMALDG"); i_ldtm = first("LDTM"); i_mma = first("UTCHMMA")
    # region boundaries: walk back from the marker to the preceding TRYWAIT-loop start is fuzzy; use markers directly
    bounds = sorted([(i_tma, "producer"), (i_ldtm, "epilogue"), (i_mma, "mma")])
    tot = sum(int(r[isamp] or 0) for r in data)
    print("== kernel %d %s samples %d" % (which, rows[s][1][22:60], tot))
    # spin loops: BRA right after TRYWAIT
    spins = []
    for i, r in enumerate(data):
        if "TRYWAIT" in src[i]:
            n = sum(int(data[j][isamp] or 0) for j in range(i, min(i + 12, len(data))) if ("BRA" in src[j] or "TRYWAIT" in src[j] or "NANOSLEEP" in src[j]))
            spins.append((i, n))
    print("   spin loops (idx, samples):", [(i, n) for i, n in spins if n > 0.005 * tot])
    print("   markers: first UTMALDG #%d, first LDTM #%d, first UTCHMMA #%d, n=%d" % (i_tma, i_ldtm, i_mma, len(data)))
