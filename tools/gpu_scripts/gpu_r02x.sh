mkdir -p gpurun_out/r02x
LIDAL_LIB=$PWD/lidal_b200/liblidal_b200_pw16.so timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -2
for lib in liblidal_b200.so liblidal_b200_pw16.so; do
LIDAL_LIB=$PWD/lidal_b200/$lib LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > gpurun_out/r02x/bench_$lib.json 2> gpurun_out/r02x/layers_$lib.txt; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02x/bench_$lib.json')); print('$lib value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
