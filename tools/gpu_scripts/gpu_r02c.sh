set -x
mkdir -p gpurun_out/r02c
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r02c/pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02c/pytest_conv.log
LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02c/bench_mode1.json 2> gpurun_out/r02c/layers_mode1.txt; echo "rc=$?"
LIDAL_PROD_MODE=0 LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02c/bench_mode0.json 2> gpurun_out/r02c/layers_mode0.txt; echo "rc=$?"
python - <<'PY'
import json
for m in (1,0):
    try:
        d=json.load(open(f'gpurun_out/r02c/bench_mode{m}.json'))
        print('mode',m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print('mode',m,'failed',e)
PY
