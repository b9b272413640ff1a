#!/bin/bash
# two epilogue groups (8 warps) + trimmed gather + single-thread MMA loop
O=gpurun_out/r02_s2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_lean.py tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_edge.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/pytest.log
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run nb1 LIDAL_NB_MAX=1
run nb2 LIDAL_NB_MAX=2
run nb1_lean0 LIDAL_NB_MAX=1 LIDAL_LEAN=0
python - <<'PY'
import json
for m in ('nb1','nb2','nb1_lean0'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2h/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
LIDAL_NB_MAX=1 LIDAL_LIB=$PWD/lidal_b200/liblidal_b200_dbg.so LIDAL_DBG=128 timeout 300 python tools/ncu_layers.py --lex 2>&1 | grep "lvl\|conv dbg" | awk '/conv dbg/{c++; if (c%12==0) print; next} {print}' > $O/segments_nb1.txt
cat $O/segments_nb1.txt
