mkdir -p gpurun_out/r02u
timeout 900 python -m pytest tests/test_gpu_pointvoxel.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not score_points and not select_regions" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --lidal-frames 300 > gpurun_out/r02u/bench.json 2> gpurun_out/r02u/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02u/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02u/bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],3), 'launches/step', d['gpu_launches']/d['steps'])
print('lidal',round(d['lidal']['value'],2),{k:round(v,1) for k,v in d['lidal']['phases_ms_max_over_ranks'].items()})
for k,v in d['roofline_by_stage'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in('ms_per_step','ms','frac','achieved')})
PY
