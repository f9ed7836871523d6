"""Does programmatic dependent launch shorten the gap between dependent kernels here? (graph replay and eager)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tools.bench_ops import timed  # noqa: E402
from di_engine_b200 import ops  # noqa: E402

L = ops.lib()
g = torch.ones(1, device='cuda')
a = torch.zeros(1024, device='cuda')
b = torch.zeros(1024, device='cuda')


def k():
    L.b200rl_scale(g.data_ptr(), a.data_ptr(), b.data_ptr(), 1024, torch.cuda.current_stream().cuda_stream)


us = timed([k] * 16, reps=50)
print(json.dumps({'pdl_env': os.environ.get('B200RL_PDL', '1'), 'tiny_kernel_us_in_graph_chain': round(us, 3)}))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(100):
        k()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(2000):
        k()
    e1.record(s)
    s.synchronize()
print(json.dumps({'pdl_env': os.environ.get('B200RL_PDL', '1'), 'tiny_kernel_us_eager_stream': round(e0.elapsed_time(e1) / 2, 3)}))
