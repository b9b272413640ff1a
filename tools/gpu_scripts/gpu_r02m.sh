mkdir -p gpurun_out/r02m
for tg in 0 1; do
LIDAL_TMA_GATHER=$tg LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02m/bench_tg$tg.json 2> gpurun_out/r02m/layers_tg$tg.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02m/bench_tg$tg.json')); print('tma_gather=$tg value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
