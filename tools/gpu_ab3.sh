#!/bin/bash
# same-box A/B of the vtws loader variants (config E): B200RL_VT_LOADER = number of leading stages copied cheaply
one() { name=$1; shift; timeout 150 python bench.py "$@" 2>/dev/null | tail -1 > gpurun_out/${TAG:-ab3}_${name}.json; }
for m in 1073741824 1 2 4 0 1073741824 1 2; do B200RL_VT_LOADER=$m one E_m${m}_$RANDOM --config E --steps 2000; done
