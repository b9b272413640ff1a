#!/bin/bash
O=gpurun_out/r02_s2z; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extras --no-lidal > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?"
}
run base
run nodn LIDAL_SORT_DN=0
run base2
run nodn2 LIDAL_SORT_DN=0
python - <<'PY'
import json
for m in ('base','nodn','base2','nodn2'):
    try:
        d=json.load(open(f'gpurun_out/r02_s2z/bench_{m}.json'))
        print(m,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'conv ms',round(d['roofline']['kernel_ms_per_step'],3),'launches',d['gpu_launches'], 'map_build', round(d['roofline_by_stage']['map_build']['ms_per_step'],3))
    except Exception as e: print(m,'failed',e)
PY
