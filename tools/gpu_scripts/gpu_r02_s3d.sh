#!/bin/bash
O=gpurun_out/r02_s3d; mkdir -p $O
LIDAL_ZSTS=1 timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run z0 LIDAL_ZSTS=0
run z1 LIDAL_ZSTS=1
python - <<'PY'
import json
for m in ('z0','z1'):
    try:
        d=json.load(open(f'gpurun_out/r02_s3d/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
paste <(awk '{print $1,$2,$3,$4,$7}' $O/layers_z0.txt) <(awk '{print $7}' $O/layers_z1.txt) | head -60
