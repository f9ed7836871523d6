#!/bin/bash
# one gpurun call: ncu --set full captures of the two one-launch step kernels + launch lists of the config D / E steps
tag=${1:-cap}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 4 -c 1 -f"
timeout 300 $NCU -k regex:gae_ppo_ws -o gpurun_out/${tag}_colws python tools/prof_step.py D 8 > gpurun_out/${tag}_colws.log 2>&1
timeout 300 $NCU -k regex:vtrace_ws -o gpurun_out/${tag}_vtws python tools/prof_step.py E 8 > gpurun_out/${tag}_vtws.log 2>&1
for c in D E P; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_$c.csv python tools/prof_step.py $c 8 > gpurun_out/${tag}_launches_$c.log 2>&1
done
timeout 300 python tools/host_overhead.py > gpurun_out/${tag}_host_overhead.json 2> gpurun_out/${tag}_host_overhead.err
ls -la gpurun_out/${tag}_*
