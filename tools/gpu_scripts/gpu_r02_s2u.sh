#!/bin/bash
O=gpurun_out/r02_s2u; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_score.py -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_a.json 2> $O/bench_a.err; echo "a rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-extras --no-lidal > $O/bench_b.json 2> $O/bench_b.err; echo "b rc=$?"
python - <<'PY'
import json
for i in ('a','b'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2u/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'host median', round(h['value_step_ms_median'],2), h['value_worst_step'], 'conv', round(d['roofline']['kernel_ms_per_step'],3), 'launches', d['gpu_launches'])
        if d.get('lidal'): print(json.dumps(d['lidal'])[:900])
    except Exception as e: print(i,'failed',e)
PY
tail -5 $O/bench_a.err
