mkdir -p gpurun_out/r02i
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu -s > gpurun_out/r02i/pytest.log 2>&1; echo "pytest rc=$?"
grep -n "vs fp32 oracle\|passed\|failed\|Error" gpurun_out/r02i/pytest.log | head
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02i/train1.json 2> gpurun_out/r02i/train1.err; echo "train rc=$?"; tail -3 gpurun_out/r02i/train1.err
python -c "
import json; d=json.load(open('gpurun_out/r02i/train1.json')); print({k:d[k] for k in ('value','ms_per_step','loss','gpu_launches','replicas_identical')})"
