#!/bin/bash
# same-box A/B of the colws loader variants (config D)
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>/dev/null | tail -1 > gpurun_out/${TAG:-ab2}_${name}.json; }
timeout 300 python -m pytest tests -m gpu -q -x -W ignore -k 'fused or gae_ppo or one_launch or smoke' 2>&1 | tail -2
one D_default --steps 2000
B200RL_COL_LOADER=1 one D_flat --steps 2000
one D_default2 --steps 2000
B200RL_COL_LOADER=1 one D_flat2 --steps 2000
one D_default_20 --steps 20 --warmup 3
B200RL_COL_LOADER=1 one D_flat_20 --steps 20 --warmup 3
