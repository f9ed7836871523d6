#!/bin/bash
# reduced 8-GPU run: config D weak scaling at the driver's K = 20 and at K = 2000, N = 1 / 2 / 4 / 8 on the same box
tag=${1:-n8mini}
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 200 python bench.py --gpus 1 "$@" 2>/dev/null | tail -1 > gpurun_out/${tag}_${name}.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" 2>/dev/null | tail -1 > gpurun_out/${tag}_${name}.json; fi; }
run n1_20 1 --steps 20 --warmup 3
run n8_20 8 --steps 20 --warmup 3
run n8_2000 8 --steps 2000
run n4_20 4 --steps 20 --warmup 3
run n2_20 2 --steps 20 --warmup 3
run n8_E 8 --config E --steps 2000
ls gpurun_out/${tag}_*
