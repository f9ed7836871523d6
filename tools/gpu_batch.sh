#!/bin/bash
# one gpurun call: GPU test suite + the bench lines of every config (outputs under gpurun_out/<tag>_*.json)
tag=${1:-run}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -W ignore 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.txt
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json; }
one 20 --steps 20 --warmup 3
one 2000 --steps 2000
one 2000_wide --steps 2000 --e2e-wide
for c in P B C E; do one $c --config $c --steps 2000; done
B200RL_VT_RES=0 one E_stream --config E --steps 2000
B200RL_VT_RES=4 one E_res4 --config E --steps 2000
one E20 --config E --steps 20 --warmup 3
