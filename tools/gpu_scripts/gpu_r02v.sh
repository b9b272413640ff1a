mkdir -p gpurun_out/r02v
for mc in 0 64 128; do
LIDAL_STAGED_MIN_COUT=$mc LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02v/bench_mc$mc.json 2> gpurun_out/r02v/layers_mc$mc.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02v/bench_mc$mc.json')); print('staged_min_cout=$mc value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
