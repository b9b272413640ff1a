mkdir -p gpurun_out/r02o
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -2
for ww in 0 1; do
LIDAL_WEIGHT_WARP=$ww LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02o/bench_ww$ww.json 2> gpurun_out/r02o/layers_ww$ww.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02o/bench_ww$ww.json')); print('weight_warp=$ww value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
