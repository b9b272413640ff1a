mkdir -p gpurun_out/r02p
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --lidal-frames 200 > gpurun_out/r02p/bench2.json 2> gpurun_out/r02p/bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r02p/bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02p/bench2.json').read().strip().splitlines()[-1])
    print('N=2 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
    print(json.dumps(d.get('lidal'),indent=0)[:1800])
except Exception as e: print('parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 10 --warmup 3 > gpurun_out/r02p/train2.json 2> gpurun_out/r02p/train2.err; echo "train2 rc=$?"
tail -3 gpurun_out/r02p/train2.err; cat gpurun_out/r02p/train2.json | tail -1 | cut -c1-600
