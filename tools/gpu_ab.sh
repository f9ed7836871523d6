#!/bin/bash
# A/B of the loader rewrite: config D and E bench lines with the flat loop / the lane-owns-a-column copies
tag=${1:-ab}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -W ignore 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; }
one D2000 --steps 2000
B200RL_COL_LOADER=1 one D2000_flat --steps 2000
one D20 --steps 20 --warmup 3
one E_res --config E --steps 2000
B200RL_VT_RES=0 one E_stream --config E --steps 2000
one P --config P --steps 2000
