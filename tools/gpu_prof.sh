#!/bin/bash
# one gpurun call: ncu --set full captures of the step kernels (outputs under gpurun_out/<tag>_*.ncu-rep)
tag=${1:-prof}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 4 -c 1 -f"
timeout 300 $NCU -k regex:vtrace_res -o gpurun_out/${tag}_vtres python tools/prof_step.py E 8 > gpurun_out/${tag}_vtres.log 2>&1
B200RL_VT_RES=0 timeout 300 $NCU -k regex:vtrace_ws -o gpurun_out/${tag}_vtws python tools/prof_step.py E 8 > gpurun_out/${tag}_vtws.log 2>&1
timeout 300 $NCU -k regex:gae_ppo_ws -o gpurun_out/${tag}_colws python tools/prof_step.py D 8 > gpurun_out/${tag}_colws.log 2>&1
ls -la gpurun_out/${tag}_*
