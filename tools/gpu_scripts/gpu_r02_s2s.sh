#!/bin/bash
O=gpurun_out/r02_s2s; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"
}
run a
run b
run prio0 LIDAL_PREP_PRIORITY=0
python - <<'PY'
import json
for i in ('a','b','prio0'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2s/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'host median', round(h['value_step_ms_median'],2), h['value_worst_step']['prepare_forward_retire_ms'], 'conv', round(d['roofline']['kernel_ms_per_step'],3), 'launches', d['gpu_launches'])
    except Exception as e: print(i,'failed',e)
PY
