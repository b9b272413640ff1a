#!/bin/bash
# trimmed gather + single-thread MMA loop: parity, then A/B over blocks per stage
O=gpurun_out/r02_s2f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/pytest.log
LIDAL_NB_MAX=2 timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py -x -q -m gpu > $O/pytest_nb2.log 2>&1; echo "pytest nb2 rc=$?"
tail -2 $O/pytest_nb2.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run nb1 LIDAL_NB_MAX=1
run nb2 LIDAL_NB_MAX=2
run nb3 LIDAL_NB_MAX=3
run nb4 LIDAL_NB_MAX=4
python - <<'PY'
import json
for m in ('nb1','nb2','nb3','nb4'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2f/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
for nb in 1 2; do
LIDAL_NB_MAX=$nb LIDAL_LIB=$PWD/lidal_b200/liblidal_b200_dbg.so LIDAL_DBG=128 timeout 300 python tools/ncu_layers.py --lex 2>&1 | grep "lvl\|conv dbg" | awk '/conv dbg/{c++; if (c%12==0) print; next} {print}' > $O/latency_nb$nb.txt
cat $O/latency_nb$nb.txt
done
