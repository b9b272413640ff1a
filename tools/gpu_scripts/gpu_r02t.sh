mkdir -p gpurun_out/r02t
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not engine and not voxelizer and not kernel_map" 2>&1 | tail -3
timeout 300 python tools/time_score.py 2>&1 | tail -1
