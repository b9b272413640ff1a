#!/bin/bash
# is the 2x step-time outlier gone with the collector parked?  8 short bench processes, 4 with LIDAL_BENCH_GC=1 (collector left on)
O=gpurun_out/r02_s2n; mkdir -p $O
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$i.json 2> $O/bench_$i.err
done
python - <<'PY'
import json
for i in range(1,9):
    try:
        d=json.load(open(f'gpurun_out/r02_s2n/bench_{i}.json'))
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), {k:round(v,2) for k,v in d['host_loop'].items() if k!='note'})
    except Exception as e: print(i,'failed',e)
PY
