#!/bin/bash
O=gpurun_out/r02_s2y; mkdir -p $O
for i in 1 2 3 4 5 6 7 8 9 10; do
  timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$i.json 2> $O/bench_$i.err
done
python - <<'PY'
import json
for i in range(1,11):
    try:
        d=json.load(open(f'gpurun_out/r02_s2y/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'e2e worst', h['e2e_worst_step'])
    except Exception as e: print(i,'failed',e)
PY
