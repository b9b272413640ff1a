mkdir -p gpurun_out/r02z
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02z/bench_full.json 2> gpurun_out/r02z/bench_full.err ) 2> gpurun_out/r02z/time_full.txt; echo "rc=$?"; tail -3 gpurun_out/r02z/time_full.txt; tail -3 gpurun_out/r02z/bench_full.err
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02z/bench_ref.json 2> gpurun_out/r02z/bench_ref.err ) 2> gpurun_out/r02z/time_ref.txt; echo "rc=$?"; tail -3 gpurun_out/r02z/time_ref.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z/bench_full.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'launches/step',d['gpu_launches']/d['steps'])
print('lidal',round(d['lidal']['value'],2),{k:round(v,1) for k,v in d['lidal']['phases_ms_max_over_ranks'].items()}, d['lidal']['selected'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('logits_rel_l2_gpu_vs_oracle'))
r=json.load(open('gpurun_out/r02z/bench_ref.json')); print('ref', r['value'], r['config']==d['config'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
