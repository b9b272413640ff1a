#!/bin/bash
# L2 prefetch warp: parity + A/B
O=gpurun_out/r02_s3b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_edge.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run off LIDAL_L2_PREFETCH_MB=0
run on24
run on1 LIDAL_L2_PREFETCH_MB=1
python - <<'PY'
import json
for m in ('off','on24','on1'):
    try:
        d=json.load(open(f'gpurun_out/r02_s3b/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
paste <(awk '{print $1,$2,$3,$4,$7}' $O/layers_off.txt) <(awk '{print $7}' $O/layers_on24.txt) <(awk '{print $7}' $O/layers_on1.txt) | head -60
