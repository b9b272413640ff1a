mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_score.py tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/r02f/pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02f/pytest.log
timeout 600 python tools/lidal_profile.py 40 > gpurun_out/r02f/lidal_profile.txt 2>&1; echo "profile rc=$?"
cat gpurun_out/r02f/lidal_profile.txt | grep -v Warning | head -60
timeout 300 python tools/time_score.py > gpurun_out/r02f/time_score.txt 2>&1; cat gpurun_out/r02f/time_score.txt | tail -2
