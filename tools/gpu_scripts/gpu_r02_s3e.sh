#!/bin/bash
O=gpurun_out/r02_s3e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_map.py tests/test_gpu_engine.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for i in a b; do timeout 300 python bench.py --no-cpu-baseline --no-extras --no-lidal > $O/bench_$i.json 2> $O/bench_$i.err; done
python - <<'PY'
import json
for m in ('a','b'):
    d=json.load(open(f'gpurun_out/r02_s3e/bench_{m}.json'))
    print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv ms',round(d['roofline']['kernel_ms_per_step'],3), d['host_loop']['value_worst_step']['prepare_forward_retire_ms'])
PY
