"""Round-2 profile summaries from one GPU visit's ncu output (tools/gpu_r2_prof.sh TAG):
   gpurun_out/launches_world_TAG.csv, launches_vae_TAG.csv (metric lists), prof_world_TAG.ncu-rep (--set full), gemm_log_vae_TAG.log
-> profiles/r02_launches_world.{md,csv}, profiles/r02_launches_vae.{md,csv}, profiles/r02_ncu_full_world.md, profiles/traffic.json
Usage: python tools/profile_r2.py TAG"""
import csv, json, os, re, shutil, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
B = 65536
WORLD = ["fwd L0   [s_t|a_t].W0^T (K=245, N=1024) +bias+ReLU+mask", "fwd L1   h0.W1^T (K=1024, N=1024) +bias+ReLU+mask",
         "fwd L2   h1.W2^T (K=1024, N=197) +bias, MSE, dLoss, db2", "wgrad L2 h1^T.g2 (M=1024, N=197, K=65536)",
         "dgrad L1 g2.W2 (K=197, N=1024) * ReLU mask, db1", "wgrad L1 h0^T.g1 (1024x1024, K=65536)",
         "dgrad L0 g1.W1 (K=1024, N=1024) * ReLU mask, db0", "wgrad L0 [s_t|a_t]^T.g0 (M=245, N=1024, K=65536)"]
WFLOP = [2.0 * B * 242 * 1024, 2.0 * B * 1024 * 1024, 2.0 * B * 1024 * 197, 2.0 * B * 1024 * 197, 2.0 * B * 1024 * 197, 2.0 * B * 1024 * 1024,
         2.0 * B * 1024 * 1024, 2.0 * B * 242 * 1024]
VAE = ["TE fwd L0 (394 -> 256)", "TE fwd L1 (256 -> 256)", "TE fwd L2 -> mu, logvar (fp32, N=64)", "MD fwd L0 (229 -> 512)", "MD fwd L1 (512 -> 512)",
       "MD fwd L2 (512 -> 512)", "MD fwd L3 + action MSE (N=45)", "WM fwd L0 (frozen, 242 -> 1024)", "WM fwd L1 (1024 -> 1024)", "WM fwd L2 + cycle MSE (N=197)",
       "WM dgrad L1 (K=197 -> 1024)", "WM dgrad L0 (1024 -> 1024)", "WM dgrad -> d a_hat (+ action-loss gradient, bias grad; N=45)", "MD wgrad L3",
       "MD dgrad L2 (K=45 -> 512)", "MD wgrad L2", "MD dgrad L1 (512 -> 512)", "MD wgrad L1", "MD dgrad L0 (512 -> 512)", "MD wgrad L0 (two M segments)",
       "MD dgrad -> dz (fp32, N=32)", "TE wgrad L2", "TE dgrad L1 (K=64 -> 256)", "TE wgrad L1", "TE dgrad L0 (256 -> 256)", "TE wgrad L0 (two M segments)"]


def rows_of(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    H = rows[0]; ix = {n: i for i, n in enumerate(H)}
    by = defaultdict(dict)
    for r in rows[1:]:
        by[int(r[ix["ID"]])]["kernel"] = r[ix["Kernel Name"]]
        by[int(r[ix["ID"]])][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    return [by[k] for k in sorted(by)]


def step_of(rows, first_pat, n):
    """the first complete step of n GEMM launches: starts at a launch matching first_pat that follows a weight-gradient launch"""
    gem = [r for r in rows if "pvae_gemm_kernel" in r["kernel"]]
    for k in range(1, len(gem)):
        if first_pat in gem[k]["kernel"] and "<3, 0, 0" in gem[k - 1]["kernel"]:
            # steady-state steps repeat the same launches: a capture that starts mid-step is completed cyclically from the step before
            return [gem[k + i] if k + i < len(gem) else gem[k + i - n] for i in range(n)]
    return gem[:n]


def table(step, labels, flops=None):
    out = ["| # | what | kernel <epi, act, tma, ctas, lean> | time us | share | tensor-pipe active | DRAM rd MB | DRAM wr MB |" + (" algorithmic TFLOP/s |" if flops else ""),
           "|---|---|---|---|---|---|---|---|" + ("---|" if flops else "")]
    tot = sum(r["gpu__time_duration.sum"] for r in step) / 1e3
    for i, (r, lab) in enumerate(zip(step, labels)):
        t = r["gpu__time_duration.sum"] / 1e3
        k = re.search(r"<[^>]*>", r["kernel"]).group(0)
        line = "| %d | %s | `%s` | %.1f | %.1f%% | %.1f%% | %.0f | %.0f |" % (i + 1, lab, k, t, 100 * t / tot, r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0),
                                                                     r.get("dram__bytes_read.sum", 0) / 1e6, r.get("dram__bytes_write.sum", 0) / 1e6)
        if flops:
            line += " %.0f |" % (flops[i] / (t * 1e-6) / 1e12)
        out.append(line)
    return "\n".join(out), tot


CMD = "ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:pvae"
tj_path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
for phase, labels, pat, n, flops in (("world", WORLD, "<0, 1, 1", 8, WFLOP), ("vae", VAE, "<0, 1, 1", 26, None)):
    src = os.path.join(ROOT, "gpurun_out", "launches_%s_%s.csv" % (phase, tag))
    if not os.path.exists(src):
        continue
    shutil.copyfile(src, os.path.join(ROOT, "profiles", "r02_launches_%s.csv" % phase))
    step = step_of(rows_of(src), pat, n)
    tab, tot = table(step, labels, flops)
    dram = sum(r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0) for r in step)
    tj["default/%s/%d" % (phase, B)] = {"dram_bytes_per_launch": dram / n, "dram_bytes_per_step": dram, "launches": n,
                                        "source": "profiles/r02_launches_%s.csv (ncu, dram__bytes_read.sum + dram__bytes_write.sum of the %d GEMM launches of one step)" % (phase, n)}
    with open(os.path.join(ROOT, "profiles", "r02_launches_%s.md" % phase), "w") as f:
        f.write("# Round 2 -- launch list of one %s-phase training step (B = 65536, dsb 197, da 45, z 32, bf16)\n\n" % ("world-model" if phase == "world" else "VAE"))
        f.write("`%s ... python bench.py --steps 2 --warmup 1 --only-phase --sustained-seconds 0 --no-cpu-baseline --phase %s` on a B200 (`tools/gpu_r2_prof.sh`); "
                "the launches are nodes of the product's captured step graph.  Serialised, cold-cache times (ncu flushes L2 between kernels): compare SHARES with the "
                "CUDA-event numbers of `bench.py`, not absolutes.  The fifth template argument marks the lean epilogue (pvae_gemm.cuh FAST).  Raw csv: `profiles/r02_launches_%s.csv`.\n\n" % (CMD, phase, phase))
        f.write(tab + "\n\n")
        f.write("Sum of the %d GEMM launches: %.1f us under ncu; DRAM traffic %.0f MB per step (%.1f MB per launch).\n" % (n, tot, dram / 1e6, dram / n / 1e6))
    print(phase, "sum %.1f us, dram %.0f MB" % (tot, dram / 1e6))
json.dump(tj, open(tj_path, "w"), indent=1)

rep = os.path.join(ROOT, "gpurun_out", "prof_world_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    R = list(csv.reader(raw.splitlines()))
    H = R[0]; ix = {n: i for i, n in enumerate(H)}
    want = [("gpu__time_duration.sum", "time us", 1.0), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %", 1.0),
            ("dram__bytes_read.sum", "DRAM read MB", 1.0), ("dram__bytes_write.sum", "DRAM write MB", 1.0), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM GB", 1.0),
            ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0), ("smsp__inst_executed.sum", "warp instr (M)", 1e-6), ("launch__registers_per_thread", "regs/thread", 1.0),
            ("sm__cycles_elapsed.max", "SM cycles", 1.0)]
    with open(os.path.join(ROOT, "profiles", "r02_ncu_full_world.md"), "w") as f:
        f.write("# Round 2 -- `ncu --set full` of the eight GEMM launches of one world-model step\n\n")
        f.write("`ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 16 -c 8` around `python bench.py --steps 2 --warmup 1 --only-phase "
                "--sustained-seconds 0 --no-cpu-baseline` (B200, `tools/gpu_r2_prof.sh`); cold-cache, serialised.  The `.ncu-rep` is scratch (`gpurun_out/`).\n\n")
        f.write("| launch | kernel | " + " | ".join(w[1] for w in want) + " |\n|---|---|" + "---|" * len(want) + "\n")
        for row, lab in zip(R[2:], WORLD):
            cells = []
            for name, _, sc in want:
                v = row[ix[name]] if name in ix else ""
                try:
                    cells.append("%.1f" % (float(v.replace(",", "")) * sc))
                except Exception:
                    cells.append(v)
            f.write("| %s | `%s` | %s |\n" % (lab, re.search(r"<[^>]*>", row[ix["Kernel Name"]]).group(0), " | ".join(cells)))
    print("wrote r02_ncu_full_world.md")
