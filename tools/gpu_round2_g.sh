#!/bin/bash
mkdir -p gpurun_out
echo "== tests: linear / fmt / transformer / cascade"
timeout 1200 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_parity.py -q -k "linear or fmt or costreg or cascade or stage_seam" 2>&1 | tail -6
echo "== breakdown"
timeout 300 python tools/profile_forward.py --iters 2 --breakdown > gpurun_out/r2g_breakdown.txt 2>&1; grep -E "ms_per_map|mvsf_" gpurun_out/r2g_breakdown.txt | paste - - | awk '{print $1, $4}' | head -20; tail -2 gpurun_out/r2g_breakdown.txt | head -1
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python tools/profile_forward.py --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2g_launches.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn].split('(')[0].replace('void ','').replace('mvsf::','')[:48]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v/1e3
tot=sum(v[1] for v in agg.values())
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:26]: print(f"{n:50s} {c:4d} {t:9.1f} us {100*t/tot:5.1f}%")
print('total us', round(tot,1))
PY
