"""Eager replay of a benchmark workload's step (bench.py --config D|B|C|E) for ncu captures -- no graphs, no CPU baseline.

    ncu --set full --clock-control none --import-source on -k regex:gae_ppo_ws -s 4 -c 1 -o gpurun_out/prof \
        python tools/prof_step.py D 8
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'D'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = bench.WORKLOADS[cfg]()
sets = [wl.device_step(wl.make_batch(i), 'cuda:0') for i in range(4)]
for i in range(steps):
    sets[i % 4]()
torch.cuda.synchronize()
print('ran', steps, 'steps of config', cfg)
