#!/bin/bash
# trimmed single-block gather path: parity + A/B against the multi-block path
O=gpurun_out/r02_s2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/pytest.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run nb2 LIDAL_NB_MAX=2
run nb1
run nb1_t2s3 LIDAL_T2_MIN_STAGES=3
run nb1_w4 LIDAL_T2_MIN_WAVES=4
python - <<'PY'
import json
for m in ('nb2','nb1','nb1_t2s3','nb1_w4'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2d/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
