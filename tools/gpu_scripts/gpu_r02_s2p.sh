#!/bin/bash
O=gpurun_out/r02_s2p; mkdir -p $O
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$i.json 2> $O/bench_$i.err
done
python - <<'PY'
import json
for i in range(1,13):
    try:
        d=json.load(open(f'gpurun_out/r02_s2p/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), 'max', round(h['value_step_ms_max'],1), round(h['e2e_step_ms_max'],1), h['value_worst_step'])
    except Exception as e: print(i,'failed',e)
PY
