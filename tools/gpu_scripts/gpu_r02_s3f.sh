#!/bin/bash
O=gpurun_out/r02_s3f; mkdir -p $O
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$i.json 2> $O/bench_$i.err
done
python - <<'PY'
import json
for i in range(1,9):
    try:
        d=json.load(open(f'gpurun_out/r02_s3f/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), 'worst e2e', round(h['e2e_worst_step']['ms'],1), 'mallocs', h['value_worst_step']['cudaMalloc_in_region'], h['e2e_worst_step']['cudaMalloc_in_region'])
    except Exception as e: print(i,'failed',e)
PY
