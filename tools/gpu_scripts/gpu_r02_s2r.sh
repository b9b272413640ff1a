#!/bin/bash
O=gpurun_out/r02_s2r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_engine.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-lidal --no-extras > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"
}
run prio0 LIDAL_PREP_PRIORITY=0
run prio0b LIDAL_PREP_PRIORITY=0
run prio1 LIDAL_PREP_PRIORITY=-1
run prio1b LIDAL_PREP_PRIORITY=-1
run prio5 LIDAL_PREP_PRIORITY=-5
python - <<'PY'
import json
for i in ('prio0','prio0b','prio1','prio1b','prio5'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2r/bench_{i}.json'))
        h=d['host_loop']
        print(i,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'host median', round(h['value_step_ms_median'],2), h['value_worst_step']['prepare_forward_retire_ms'], 'conv', round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(i,'failed',e)
PY
