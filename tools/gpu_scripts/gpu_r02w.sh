mkdir -p gpurun_out/r02w
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02w/bench_$i.json 2> gpurun_out/r02w/err_$i.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02w/bench_$i.json')); print('run $i value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02w/bench_50.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02w/bench_50.json')); print('50 steps: value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3))"
