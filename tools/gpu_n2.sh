run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 "$@" 2>gpurun_out/r3h_${name}.err | tail -1 > gpurun_out/r3h_${name}.json; }
run n2_D20 --steps 20 --warmup 3
run n2_P --config P --steps 2000
run n2_E --config E --steps 2000
run n2_ref --impl reference --steps 4 --warmup 1
timeout 300 python -m pytest tests/test_p2p_gpu.py -m gpu -q -x -W ignore 2>&1 | tail -2
ls gpurun_out/r3h_*
