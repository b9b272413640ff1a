mkdir -p gpurun_out/r02r
LIDAL_NB_MAX=3 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -2
for nb in 1 2 3 4; do
LIDAL_NB_MAX=$nb LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02r/bench_nb$nb.json 2> gpurun_out/r02r/layers_nb$nb.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02r/bench_nb$nb.json')); print('nb_max=$nb value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
