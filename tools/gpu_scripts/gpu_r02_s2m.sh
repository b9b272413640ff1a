#!/bin/bash
# checkpoint: full GPU suite, default bench (twice), launch list under ncu, one ncu --set full capture of the level-0 conv
O=gpurun_out/r02_s2m; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --no-cpu-baseline --no-lidal > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('bench','bench2'):
    d=json.load(open(f'gpurun_out/r02_s2m/{f}.json'))
    print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv',round(d['roofline']['kernel_ms_per_step'],3),'launches',d.get('gpu_launches'), 'lidal', d.get('lidal_frames_per_sec'))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-lidal > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 10 -c 1 -o $O/conv_l0_96 python tools/ncu_layers.py --lex > $O/ncu1.log 2>&1; tail -2 $O/ncu1.log
ls -la $O
