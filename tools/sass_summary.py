"""profiles/sass_summary.md: opcode evidence from `cuobjdump -sass` of the shipped library (run in the build container, no GPU needed).
Usage: python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "physicsvae_b200", "lib", "libpvae_sm100.so")
KEY = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "UBLKCP", "SYNCS", "ELECT", "UCGABAR", "HMMA", "LDSM",
       "FADD2", "REDG", "RED", "ATOMG", "ATOM", "F2FP", "PRMT", "MEMBAR", "ACQBULK", "LDGDEPBAR", "HGMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    ptx = subprocess.run(["cuobjdump", "-ptx", LIB], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = fn.replace("pvae::", "").replace("(pvae::GemmParams)", "").replace("void ", "")
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
            per[fn][m.group(1) + m.group(2)] += 0
            if m.group(2):
                per[fn]["full:" + m.group(1) + m.group(2)] += 1
    total = collections.Counter()
    full = collections.Counter()
    for c in per.values():
        for k, v in c.items():
            if k.startswith("full:"):
                full[k[5:]] += v
            elif "." not in k:
                total[k] += v
    arch = re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout + sass)
    print("# SASS summary of `physicsvae_b200/lib/libpvae_sm100.so`\n")
    print("`cuobjdump -sass` of the library built by `__graft_entry__.build()` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`); "
          "regenerate with `python tools/sass_summary.py > profiles/sass_summary.md`.  Target: %s.  %d kernels, %d SASS instructions.\n" % (
              ", ".join(sorted(set(arch))) or "sm_100a", len(per), sum(total.values())))
    print("## Blackwell-native opcodes (whole library)\n")
    print("| SASS opcode | count | what it is |\n|---|---|---|")
    what = {"UTCHMMA": "`tcgen05.mma.kind::f16` (5th-gen tensor core, accumulator in TMEM)", "UTCBAR": "`tcgen05.commit` -> mbarrier (`.2CTA.MULTICAST`: both CTAs of a pair)",
            "UTCATOMSWS": "`tcgen05.alloc` / `dealloc` (TMEM columns)", "LDTM": "`tcgen05.ld` (TMEM -> registers, epilogue)",
            "UTMALDG": "TMA tile load `cp.async.bulk.tensor.3d` (`.2CTA`: pair-addressed barrier)", "UTMASTG": "TMA tile store (epilogue slabs)",
            "UTMACMDFLUSH": "`cp.async.bulk.commit_group`", "UBLKCP": "`cp.async.bulk` (untiled bulk copy; `.S.G`: global -> shared, the gradient exchange pulling peer slices; `.G.S`: shared -> global)",
            "UTMAPF": "`cp.async.bulk.prefetch.tensor` (the optional L2 operand prefetch, off by default)", "SYNCS": "mbarrier init / arrive / expect_tx / try_wait", "ELECT": "`elect.sync` (one issuing lane)",
            "UCGABAR": "cluster barrier (CTA pair)", "HMMA": "`mma.sync` (legacy tensor path: only the optional PVAE_CS_MMA=1 column-sum experiment)",
            "LDSM": "`ldmatrix` (same experiment)", "FADD2": "`add.f32x2` (packed fp32 adds: bias, column sums)", "F2FP": "`cvt.rn(.relu).bf16x2.f32`",
            "PRMT": "`prmt` (ReLU mask expansion, byte sign replication)", "REDG": "`red.global.add.f32` (split-K weight gradients, bias gradients)",
            "ACQBULK": "griddepcontrol / bulk acquire", "HGMMA": "wgmma (Hopper) -- must be absent"}
    for k in KEY:
        if total.get(k) or k == "HGMMA":
            print("| `%s` | %d | %s |" % (k, total.get(k, 0), what.get(k, "")))
    print("\nVariants seen: " + ", ".join("`%s` x%d" % (k, v) for k, v in sorted(full.items()) if k.split(".")[0] in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "UTCATOMSWS", "MULTIMEM") or "MULTIMEM" in k or k.startswith("LD.") and "SYS" in k))
    mm = collections.Counter(re.findall(r"(multimem\.[a-z_]+|griddepcontrol\.[a-z_]+|tcgen05\.[a-z_]+(?:\.cta_group::\d)?|cp\.async\.bulk\.tensor\.\dd|ld\.relaxed\.sys|st\.relaxed\.sys|st\.release\.sys|ld\.acquire\.sys)", ptx))
    print("\n## PTX mnemonics (embedded PTX of the same library)\n")
    print(", ".join("`%s` x%d" % (k, v) for k, v in sorted(mm.items())) or "(no PTX embedded)")
    print("\n## Per kernel\n")
    print("| kernel | instructions | UTCHMMA | LDTM | UTMALDG | UTMASTG | HMMA |\n|---|---|---|---|---|---|---|")
    for fn, c in per.items():
        n = sum(v for k, v in c.items() if "." not in k and not k.startswith("full:"))
        print("| `%s` | %d | %d | %d | %d | %d | %d |" % (fn[:90], n, c.get("UTCHMMA", 0), c.get("LDTM", 0), c.get("UTMALDG", 0), c.get("UTMASTG", 0), c.get("HMMA", 0)))


if __name__ == "__main__":
    main()
