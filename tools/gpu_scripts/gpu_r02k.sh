mkdir -p gpurun_out/r02k
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "engine" 2>&1 | tail -3
for tg in 0 1; do
LIDAL_TMA_GATHER=$tg LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02k/bench_tg$tg.json 2> gpurun_out/r02k/layers_tg$tg.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02k/bench_tg$tg.json')); print('tma_gather=$tg value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
