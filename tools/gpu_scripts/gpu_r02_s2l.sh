#!/bin/bash
# tile-shape / locality A/B on the new kernel + cycle accounting
O=gpurun_out/r02_s2l; mkdir -p $O
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run base LIDAL_NB_MAX=2
run tile128 LIDAL_NB_MAX=2 LIDAL_TILE128=1
run waves4 LIDAL_NB_MAX=2 LIDAL_T2_MIN_WAVES=4
run waves16 LIDAL_NB_MAX=2 LIDAL_T2_MIN_WAVES=16
run chunk14 LIDAL_NB_MAX=2 LIDAL_MASK_CHUNK_SHIFT=14
run chunk16 LIDAL_NB_MAX=2 LIDAL_MASK_CHUNK_SHIFT=16
run nb3 LIDAL_NB_MAX=3
python - <<'PY'
import json
for m in ('base','tile128','waves4','waves16','chunk14','chunk16','nb3'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2l/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
LIDAL_NB_MAX=2 LIDAL_LIB=$PWD/lidal_b200/liblidal_b200_dbg.so LIDAL_DBG=128 timeout 300 python tools/ncu_layers.py --lex 2>&1 | grep "lvl\|conv dbg" | awk '/conv dbg/{c++; if (c%12==0) print; next} {print}' > $O/segments_nb2.txt
cat $O/segments_nb2.txt
