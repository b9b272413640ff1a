#!/bin/bash
N=${1:-4}
O=gpurun_out/r02_s3_${N}gpu; mkdir -p $O
nproc; nvidia-smi topo -m 2>/dev/null | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('$O/bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'n',d['n_gpus'],'lidal',d.get('lidal_frames_per_sec'))
        print(d['host_loop']['value_worst_step'], d['host_loop']['e2e_worst_step'])
        print(json.dumps(d.get('collective')))
PY
