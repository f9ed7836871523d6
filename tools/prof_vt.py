"""Eager replay of the V-trace one-launch step (config E) for ncu captures: forward launch + verification launch.

    ncu --set full ... -k regex:vtrace_ws -s 4 -c 1 -o gpurun_out/prof python tools/prof_vt.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tools.exp_vt as ev  # noqa: E402  (builds the buffer sets)

steps = 8
for i in range(steps):
    x = ev.sets[i % len(ev.sets)]
    x.fused()
    x.verify()
torch.cuda.synchronize()
print('ran', steps, 'steps')
