#!/bin/bash
# ring depth vs staged epilogue: free the staging buffers of permuted outputs (direct stores) for more stages
O=gpurun_out/r02_s2x; mkdir -p $O
run() { name=$1; shift
  env "$@" LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/layers_$name.txt; echo "$name rc=$?"
}
run min64
run min128 LIDAL_STAGED_MIN_COUT=128
run min192 LIDAL_STAGED_MIN_COUT=192
run min512 LIDAL_STAGED_MIN_COUT=512
python - <<'PY'
import json
for m in ('min64','min128','min192','min512'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2x/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv ms',round(d['roofline']['kernel_ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
paste <(awk '{print $1,$2,$3,$4,$7}' $O/layers_min64.txt) <(awk '{print $7}' $O/layers_min128.txt) <(awk '{print $7}' $O/layers_min192.txt) <(awk '{print $7}' $O/layers_min512.txt) | head -60
