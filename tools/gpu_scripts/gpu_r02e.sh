mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_score.py -x -q -m gpu > gpurun_out/r02e/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -15 gpurun_out/r02e/pytest_new.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s > gpurun_out/r02e/pytest_full.log 2>&1; echo "pytest fullsize rc=$?"
tail -15 gpurun_out/r02e/pytest_full.log
timeout 900 python bench.py --steps 10 --warmup 3 --lidal-frames 200 > gpurun_out/r02e/bench.json 2> gpurun_out/r02e/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02e/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02e/bench.json'))
    print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4))
    print(json.dumps(d.get('lidal'),indent=1)[:2500])
    print(json.dumps(d.get('cpu_baseline'),indent=1))
except Exception as e: print('bench parse failed',e)
PY
