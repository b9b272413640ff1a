mkdir -p gpurun_out/r02q
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not score_points and not select_regions" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --lidal-frames 300 > gpurun_out/r02q/bench.json 2> gpurun_out/r02q/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02q/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02q/bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],3))
print('lidal',round(d['lidal']['value'],2),{k:round(v,1) for k,v in d['lidal']['phases_ms_max_over_ranks'].items()}, d['lidal']['checks'])
print(d['cpu_baseline'].get('logits_rel_l2_gpu_vs_oracle'))
PY
