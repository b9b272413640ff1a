#!/bin/bash
# 8-GPU validation: infer+LiDAL bench and the DDP training workload, as the driver launches them.
mkdir -p gpurun_out/r02_8gpu
O=gpurun_out/r02_8gpu
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
tail -c 3000 $O/bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --workload train > $O/train_${N}gpu.json 2> $O/train_${N}gpu.err
tail -c 1500 $O/train_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/ref_${N}gpu.json 2> $O/ref_${N}gpu.err
tail -c 800 $O/ref_${N}gpu.json
