#!/bin/bash
# A/B of loader / wait variants on ONE box: config D (and E) bench lines
tag=${1:-ab}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -W ignore 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; }
one D_g8 --steps 2000
B200RL_COL_LOADER=1 one D_flat --steps 2000
B200RL_COL_LOADER=2 one D_g4 --steps 2000
B200RL_COL_WAIT_NS=200 one D_g8_w200 --steps 2000
B200RL_COL_WAIT_NS=2000 one D_g8_w2000 --steps 2000
B200RL_COL_LOADER=1 B200RL_COL_WAIT_NS=2000 one D_flat_w2000 --steps 2000
one D_g8_again --steps 2000
one D20 --steps 20 --warmup 3
one E --config E --steps 2000
one P --config P --steps 2000
