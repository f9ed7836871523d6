#!/bin/bash
# two-rank sanity run of the bench lines and the peer-memory test
tag=${1:-n2}
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; }
run n2_D20 --steps 20 --warmup 3
run n2_D2000 --steps 2000
run n2_E --config E --steps 2000
timeout 300 python -m pytest tests/test_p2p_gpu.py -m gpu -q -x -W ignore 2>&1 | tail -2
timeout 100 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${tag}_n1_D20.json
ls gpurun_out/${tag}_*
