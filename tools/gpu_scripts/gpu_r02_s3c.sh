#!/bin/bash
# register-staged gather (LDG.128 -> STS.128, software-pipelined): parity + A/B against cp.async
O=gpurun_out/r02_s3c; mkdir -p $O
LIDAL_REGSTAGE=1 timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_edge.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run cpasync LIDAL_REGSTAGE=0
run regstage LIDAL_REGSTAGE=1
python - <<'PY'
import json
for m in ('cpasync','regstage'):
    try:
        d=json.load(open(f'gpurun_out/r02_s3c/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
paste <(awk '{print $1,$2,$3,$4,$7}' $O/layers_cpasync.txt) <(awk '{print $7}' $O/layers_regstage.txt) | head -60
