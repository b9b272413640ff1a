#!/bin/bash
# lean conv producers: parity tests, then A/B of the per-layer table
O=gpurun_out/r02_s2b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py -x -q -m gpu > $O/pytest_lean.log 2>&1; echo "pytest lean rc=$?"
tail -15 $O/pytest_lean.log
LIDAL_LEAN_ARRIVE=1 timeout 900 python -m pytest tests/test_gpu_conv_lean.py -x -q -m gpu > $O/pytest_lean_arrive.log 2>&1; echo "pytest lean arrive rc=$?"
tail -5 $O/pytest_lean_arrive.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/pytest_conv.log 2>&1; echo "pytest conv/engine rc=$?"
tail -5 $O/pytest_conv.log
run() { # name, env...
  name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run lean0 LIDAL_LEAN=0
run lean3 LIDAL_LEAN=3
run lean3_arrive LIDAL_LEAN=3 LIDAL_LEAN_ARRIVE=1
run lean3_arrive_lag1 LIDAL_LEAN=3 LIDAL_LEAN_ARRIVE=1 LIDAL_LEAN_LAG=1
run lean1 LIDAL_LEAN=1
run lean2 LIDAL_LEAN=2
python - <<'PY'
import json
for m in ('lean0','lean3','lean3_arrive','lean3_arrive_lag1','lean1','lean2'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2b/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
