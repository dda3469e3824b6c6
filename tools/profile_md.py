"""Turn one GPU visit's ncu output into the committed summaries under profiles/.

  python tools/profile_md.py TAG ROUND [bench_log]
    gpurun_out/launches_TAG.csv   (ncu --metrics gpu__time_duration.sum,... launch list)   -> profiles/rROUND_launches_world.{csv,md}
    gpurun_out/prof_TAG.ncu-rep   (ncu --set full of the GEMM launches of one world step)   -> profiles/rROUND_ncu_full_world.md, traffic.json
World step at default dims: the launches of one step are labelled from their kernel instantiation and order.
"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
B, DSB, DA, H = 65536, 197, 45, 1024

# (label, algorithmic FLOPs) of the eight GEMM launches of one world step, in launch order
STEP = [
    ("fwd L0   [s_t|a_t].W0^T (K=245: one segment of the resident row, N=1024) +bias+ReLU+mask", 2.0 * B * (DSB + DA) * H),
    ("fwd L1   h0.W1^T (K=1024, N=1024) +bias+ReLU+mask", 2.0 * B * H * H),
    ("fwd L2   h1.W2^T (K=1024, N=197) +bias, MSE, dLoss, db2", 2.0 * B * H * DSB),
    ("wgrad L2 h1^T.g2 (M=1024, N=197, K=65536)", 2.0 * B * H * DSB),
    ("dgrad L1 g2.W2 (K=197, N=1024) * ReLU mask, db1", 2.0 * B * H * DSB),
    ("wgrad L1 h0^T.g1 (1024x1024, K=65536)", 2.0 * B * H * H),
    ("dgrad L0 g1.W1 (K=1024, N=1024) * ReLU mask, db0", 2.0 * B * H * H),
    ("wgrad L0 [s_t|a_t]^T.g0 (M=245, N=1024, K=65536)", 2.0 * B * (DSB + DA) * H),
]


def launch_rows(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    Hd = rows[h]
    iid, ik, im, iv, iu = Hd.index("ID"), Hd.index("Kernel Name"), Hd.index("Metric Name"), Hd.index("Metric Value"), Hd.index("Metric Unit")
    L, names = defaultdict(dict), {}
    for r in rows[h + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        L[int(r[iid])][r[im]] = v * scale
        names[int(r[iid])] = r[ik]
    return [(i, names[i], L[i]) for i in sorted(L)]


def one_step(launches):
    """the last complete world step in the list: fwd L0 is the first <0, 1, 1, *> launch after a wgrad"""
    gem = [(i, n, m) for i, n, m in launches if "pvae_gemm_kernel" in n]
    starts = [k for k in range(len(gem)) if "<0, 1, 1" in gem[k][1] and (k == 0 or "<3, 0, 0" in gem[k - 1][1])]
    starts = [k for k in starts if k + len(STEP) <= len(gem)]
    k = starts[-1]
    return gem[k:k + len(STEP)]


def md_launches():
    src = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
    dst = os.path.join(ROOT, "profiles", "r%s_launches_world.csv" % rnd)
    shutil.copyfile(src, dst)
    step = one_step(launch_rows(src))
    tot = sum(m["gpu__time_duration.sum"] for _, _, m in step)
    out = ["# Round %s -- launch list of one world-model training step (B = 65536, dsb 197, da 45, bf16)\n" % rnd,
           "Command (on the B200 box): `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,"
           "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 40 -c 30 --csv python bench.py --steps 2 --warmup 1 "
           "--no-graph --no-cpu-baseline`",
           "(per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event numbers of bench.py, not absolutes). "
           "Raw csv: `profiles/r%s_launches_world.csv`.\n" % rnd,
           "| # | launch (pvae_gemm_kernel<epilogue, act, tma, ctas-per-mma>) | what | time us | share | tensor-pipe active | DRAM rd MB | DRAM wr MB | algorithmic TFLOP/s |",
           "|---|---|---|---|---|---|---|---|---|"]
    fl = 0.0
    for k, ((i, n, m), (label, flops)) in enumerate(zip(step, STEP)):
        t = m["gpu__time_duration.sum"]
        fl += flops
        out.append("| %d | `%s` | %s | %.1f | %.1f%% | %.1f%% | %.0f | %.0f | %.0f |" % (
            k + 1, n.replace("void ", "").replace("(GemmParams)", ""), label, t, 100 * t / tot,
            m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0), m.get("dram__bytes_read.sum", 0.0),
            m.get("dram__bytes_write.sum", 0.0), flops / t / 1e6))
    rd = sum(m.get("dram__bytes_read.sum", 0.0) for _, _, m in step)
    wr = sum(m.get("dram__bytes_write.sum", 0.0) for _, _, m in step)
    out.append("\nSum of the %d GEMM launches: %.0f us (ncu, serialised) for %.1f algorithmic GFLOP -> %.0f TFLOP/s." % (len(step), tot, fl / 1e9, fl / tot / 1e6))
    out.append("DRAM traffic of the step's GEMM launches: %.0f MB read + %.0f MB written." % (rd, wr))
    if len(sys.argv) > 3:
        d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
        r = d["roofline"]
        out.append("\n`bench.py` of the same build (CUDA events, graph replay, no profiler): %.3f ms per step = %.1f M transitions/s; GEMM launch "
                   "sequence alone %.3f ms = %.0f TFLOP/s = %.1f %% of the measured sustained bf16 peak (%.0f TFLOP/s)." % (
                       d["ms_per_step"], d["value"] / 1e6, r["kernel_ms_per_step"], r["achieved"], 100 * r["frac"], r["peak"]))
    open(os.path.join(ROOT, "profiles", "r%s_launches_world.md" % rnd), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def md_full():
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    Hd = rows[0]
    col = lambda n: Hd.index(n)
    want = [("time us", "gpu__time_duration.sum"), ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM write MB", "dram__bytes_write.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
            ("issue slots busy %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), ("warp instr (M)", "smsp__inst_executed.sum"),
            ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
    data = rows[2:][:len(STEP)]     # consecutive launches = one step's worth, possibly starting mid-step
    # rotate so that the table starts at fwd L0
    k0 = [k for k, r in enumerate(data) if "<0, 1, 1" in r[col("Kernel Name")]][0]
    data = data[k0:] + data[:k0]
    out = ["# Round %s -- `ncu --set full` of the GEMM launches of one world-model step\n" % rnd,
           "Workload: `python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline` (world phase, dsb 197 / da 45, B = 65536, bf16), captured with",
           "`ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 27 -c 9` on a B200 (`gpurun`); numbers under the profiler are",
           "cold-cache and serialised. The `.ncu-rep` itself is scratch (`gpurun_out/prof_%s.ncu-rep`), this file is its summary.\n" % tag,
           "| launch | kernel | " + " | ".join(w[0] for w in want) + " |", "|---|---|" + "---|" * len(want)]
    tot = 0.0
    for r, (label, _) in zip(data, STEP):
        cells = []
        for name, key in want:
            v = float(r[col(key)].replace(",", ""))
            if key == "smsp__inst_executed.sum":
                v /= 1e6
            cells.append(("%.1f" % v) if v != int(v) or "%" in name or "us" in name else "%d" % v)
        tot += float(r[col("dram__bytes_read.sum")]) + float(r[col("dram__bytes_write.sum")])
        out.append("| %s | `%s` | %s |" % (label,
                                           r[col("Kernel Name")].replace("void ", "").replace("(GemmParams)", ""), " | ".join(cells)))
    n = min(len(data), len(STEP))
    out.append("\nDRAM traffic of these %d launches: %.0f MB (`dram__bytes_read.sum + dram__bytes_write.sum`), %.0f MB per launch -> `profiles/traffic.json`"
               " (`roofline.traffic` of bench.py)." % (n, tot, tot / n))
    open(os.path.join(ROOT, "profiles", "r%s_ncu_full_world.md" % rnd), "w").write("\n".join(out) + "\n")
    tj = {"default/world/65536": {"dram_bytes_per_launch": tot * 1e6 / n, "dram_bytes_per_step": tot * 1e6, "launches": n,
                                   "source": "profiles/r%s_ncu_full_world.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % rnd}}
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("\n".join(out))


md_launches()
md_full()
