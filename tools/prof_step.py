"""Eager replay of the benchmark's learner step (config D) for ncu captures -- no graphs, no CPU baseline.

    ncu --set full ... -k regex:"gae_tile|ppo_fwd|ppo_bwd" -s 12 -c 3 -o gpurun_out/prof python tools/prof_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else 'onepass'
mode = {'onepass': 'onepass', 'three': True, 'unfused': False}[mode]
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused=mode) for i in range(4)]
for i in range(steps):
    sets[i % 4]()
torch.cuda.synchronize()
print('ran', steps, 'steps')
