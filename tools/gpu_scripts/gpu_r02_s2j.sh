#!/bin/bash
# short prologue (fastpro): parity + A/B
O=gpurun_out/r02_s2j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest.log
LIDAL_LEAN=3 timeout 900 python -m pytest tests/test_gpu_conv_lean.py -x -q -m gpu > $O/pytest_lean3.log 2>&1; echo "pytest lean3 rc=$?"
tail -2 $O/pytest_lean3.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run fast0 LIDAL_FASTPRO=0
run fast1 LIDAL_FASTPRO=1
run fast1_nb2 LIDAL_FASTPRO=1 LIDAL_NB_MAX=2
python - <<'PY'
import json
for m in ('fast0','fast1','fast1_nb2'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2j/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
