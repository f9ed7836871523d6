#!/bin/bash
# reduced 8-GPU run: config D weak scaling at the driver's K = 20 and at K = 2000, N = 1 and 8 on the same box
tag=${1:-n8mini}
mkdir -p gpurun_out
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${tag}_n1_20.json
run() { name=$1; shift; timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 "$@" 2>/dev/null | tail -1 > gpurun_out/${tag}_${name}.json; }
run n8_20 --steps 20 --warmup 3
run n8_2000 --steps 2000
ls gpurun_out/${tag}_*
