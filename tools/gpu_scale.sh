#!/bin/bash
# one gpurun --gpus N call: the scaling lines of configs D (weak K=20 / K=2000, strong) and E, plus N=1 on the same box
N=${1:-8}
tag=${2:-scale}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
run() { name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 200 python bench.py --gpus 1 "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; fi; }
run n1_20 1 --steps 20 --warmup 3
run n${N}_20 $N --steps 20 --warmup 3
run n${N}_2000 $N --steps 2000
run n${N}_strong $N --steps 2000 --scaling strong
run n${N}_E $N --config E --steps 2000
run n${N}_E_strong $N --config E --steps 2000 --scaling strong
run n${N}_nccl $N --steps 2000 --collective nccl
run n${N}_p2pk $N --steps 2000 --collective p2p-kernel
if [ "$N" -ge 4 ]; then run n2_20 2 --steps 20 --warmup 3; run n4_20 4 --steps 20 --warmup 3; fi
timeout 200 python -m pytest tests/test_p2p_gpu.py -m gpu -q -x -W ignore 2>&1 | tail -3 > gpurun_out/${tag}_p2p_pytest.txt
ls gpurun_out/${tag}_* | head -40
