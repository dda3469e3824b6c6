#!/bin/bash
# epilogue timing experiments: kernel durations of one world step with parts of the TMA epilogue skipped (PVAE_DBG bits)
# needs the debug build (-DPVAE_DEBUG_HOOKS -> physicsvae_b200/lib/libpvae_sm100_dbg.so), see tools/README.md
mkdir -p gpurun_out
for d in 0 1 2 4 8 16 31; do
PVAE_LIB=$PWD/physicsvae_b200/lib/libpvae_sm100_dbg.so PVAE_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pvae_gemm -s 24 -c 8 --csv --log-file gpurun_out/dbg_$d.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/dbg_$d.log 2>&1
echo "dbg=$d: $(python - <<P
import csv
rows=list(csv.reader(open('gpurun_out/dbg_$d.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[h]; iv=H.index('Metric Value'); iu=H.index('Metric Unit')
vals=[float(r[iv].replace(',',''))/(1e3 if r[iu] in ('ns','nsecond') else 1) for r in rows[h+1:] if len(r)>iv]
print(' '.join('%.1f'%v for v in vals), ' total %.1f'%sum(vals))
P
)"
done
