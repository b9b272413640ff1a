#!/bin/bash
# session-2 baseline: full GPU test suite, default bench line, per-layer conv table
O=gpurun_out/r02_s2a; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -3 $O/bench.err
LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_layers.json 2> $O/layers.txt; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s2a/bench.json'))
print('value',round(d['value'],1),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'launches',d.get('gpu_launches'))
print(json.dumps(d.get('lidal'))[:1500])
print(json.dumps(d.get('clocks')))
PY
