#!/bin/bash
# same-box A/B of the colws loader variants (config D): B200RL_COL_LOADER 0 = first stage cheap, 5 / 6 / 7 = first 2 / 3 / 4 stages
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>/dev/null | tail -1 > gpurun_out/${TAG:-ab2}_${name}.json; }
for m in 0 5 6 0 5 6; do B200RL_COL_LOADER=$m one D_m${m}_$RANDOM --steps 2000; done
