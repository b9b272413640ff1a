#!/bin/bash
# N-GPU validation as the driver launches it: infer + LiDAL bench, then the DDP training workload (short)
N=${1:-2}
O=gpurun_out/r02_s2_${N}gpu; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -3 $O/bench.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload train > $O/train.json 2> $O/train.err; echo "train rc=$?"
tail -3 $O/train.err
python - <<PY
import json
d=json.load(open('$O/bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'n',d['n_gpus'],'lidal',d.get('lidal_frames_per_sec'))
print(json.dumps(d.get('collective')))
print(json.dumps((d.get('lidal') or {}).get('phases_ms_max_over_ranks')))
try:
    t=json.load(open('$O/train.json')); print('train', {k:t[k] for k in ('metric','value','unit','ms_per_step','n_gpus') if k in t})
except Exception as e: print('train parse failed', e)
PY
