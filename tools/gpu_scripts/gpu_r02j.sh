mkdir -p gpurun_out/r02j
for pm in 0 2; do
echo "== PROD_MODE=$pm"; LIDAL_PROD_MODE=$pm timeout 120 python tools/ncu_layers.py --lex 2>&1 | grep lvl
done > gpurun_out/r02j/modes.txt 2>&1
cat gpurun_out/r02j/modes.txt
LIDAL_PROD_MODE=2 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -3
LIDAL_PROD_MODE=2 LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02j/bench_mode2.json 2> gpurun_out/r02j/layers_mode2.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02j/bench_mode2.json')); print('mode2 value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
