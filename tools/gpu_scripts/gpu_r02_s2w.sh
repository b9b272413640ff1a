#!/bin/bash
# producers are now THE bottleneck (MMA warp waits 93 %, LSU queue full 27 % of producer samples): does the TMA gather win now?
O=gpurun_out/r02_s2w; mkdir -p $O
LIDAL_TMA_GATHER=2 LIDAL_FASTPRO=0 timeout 600 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu > $O/pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -3 $O/pytest_tma.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run base
run tma1 LIDAL_TMA_GATHER=1 LIDAL_FASTPRO=0
run tma2 LIDAL_TMA_GATHER=2 LIDAL_FASTPRO=0
run lean3 LIDAL_LEAN=3
python - <<'PY'
import json
for m in ('base','tma1','tma2','lean3'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2w/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
