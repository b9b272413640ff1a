#!/bin/bash
O=gpurun_out/r02_s3a; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"
}
run a
run b
python - <<'PY'
import json
for m in ('a','b'):
    try:
        d=json.load(open(f'gpurun_out/r02_s3a/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv ms',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'launches',d['gpu_launches'], 'map_build', round(d['roofline_by_stage']['map_build']['ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
