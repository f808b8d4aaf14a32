#!/bin/bash
mkdir -p gpurun_out
echo "== staging microbenchmark"
timeout 120 tools/stage_microbench.bin > gpurun_out/r2_stage_microbench.jsonl 2>&1; cat gpurun_out/r2_stage_microbench.jsonl
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py tests/test_gpu_boundary.py -q -k "linear or fmt or costreg or cascade or stage_seam or install or batch" 2>&1 | tail -5
echo "== breakdown"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2i_breakdown.txt 2>&1; grep -E "costreg_tr|fmt_forward" -A1 gpurun_out/r2i_breakdown.txt | grep -E "mvsf|ms_per"; grep "total ms" gpurun_out/r2i_breakdown.txt
echo "== launch list (linear)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches.csv python tools/profile_forward.py --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2i_launches.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0].replace('void ','').replace('mvsf::','')[:48]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v/1e3
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    if 'at::' in n: continue
    print(f"{n:50s} {c:4d} {t:9.1f} us")
PY
