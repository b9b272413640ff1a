#!/bin/bash
O=gpurun_out/r02_s2v; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-extras --no-lidal > $O/bench_b.json 2> $O/bench_b.err; echo "b rc=$?"
python - <<'PY'
import json
for i in ('bench','bench_b'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2v/{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'host median', round(h['value_step_ms_median'],2), h['value_worst_step'], 'conv', round(d['roofline']['kernel_ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'], 'lidal', d.get('lidal_frames_per_sec'))
    except Exception as e: print(i,'failed',e)
PY
