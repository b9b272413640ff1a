#!/bin/bash
# final evidence of round 2: full GPU suite, smoke(), default bench line, per-layer table, reference arm, launch list under ncu,
# ncu --set full captures of three conv shapes
O=gpurun_out/r02_final; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
LIDAL_LAYER_TABLE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-lidal > $O/bench_layers.json 2> $O/layers.txt; echo "layers rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_final/bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],4),'conv',round(d['roofline']['kernel_ms_per_step'],3),'launches',d.get('gpu_launches'),'lidal',d.get('lidal_frames_per_sec'))
print(json.dumps(d.get('clocks')), json.dumps(d.get('cpu_baseline'))[:300])
r=json.load(open('gpurun_out/r02_final/bench_reference.json')); print('reference', r.get('value'), r.get('unit'), r.get('impl'))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-lidal > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv_tc -s 10 -c 1 -o $O/conv_l0_96 python tools/ncu_layers.py --lex > $O/ncu_conv1.log 2>&1
$NCU -k regex:conv_tc -s 46 -c 1 -o $O/conv_l2_128 python tools/ncu_layers.py --lex > $O/ncu_conv2.log 2>&1
$NCU -k regex:conv_tc -s 58 -c 1 -o $O/conv_l3_256 python tools/ncu_layers.py --lex > $O/ncu_conv3.log 2>&1
ls -la $O
