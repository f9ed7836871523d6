"""Calibration: what a plain device copy achieves at the byte counts of this path (graph replay, rotated buffers)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tools.bench_ops import timed  # noqa: E402

for mb in (1, 4, 6.3, 12.6, 27, 54, 128, 512):
    n = int(mb * 1e6 / 8)  # read n floats + write n floats = mb MB of traffic
    nsets = max(2, int(400e6 / (mb * 1e6)) + 1)
    nsets = min(nsets, 64)
    srcs = [torch.randn(n, device='cuda') for _ in range(nsets)]
    dsts = [torch.empty(n, device='cuda') for _ in range(nsets)]
    us = timed([lambda s=s, d=d: d.copy_(s) for s, d in zip(srcs, dsts)], reps=20)
    print(json.dumps({'traffic_MB': mb, 'us': round(us, 2), 'GBps': round(mb * 1e3 / us, 1)}))
    del srcs, dsts
from di_engine_b200 import ops  # noqa: E402
L = ops.lib()
for mb in (6.3, 12.6, 27, 54.5, 67, 128):
    n = int(mb * 1e6 / 8) // 4 * 4
    nsets = min(64, max(2, int(400e6 / (mb * 1e6)) + 1))
    srcs = [torch.randn(n, device='cuda') for _ in range(nsets)]
    dsts = [torch.empty(n, device='cuda') for _ in range(nsets)]
    for cps in (2, 4, 8):
        us = timed([lambda s=s, d=d, cps=cps: L.b200rl_probe_copy(s.data_ptr(), d.data_ptr(), n, cps,
                                                                    torch.cuda.current_stream().cuda_stream)
                    for s, d in zip(srcs, dsts)], reps=20)
        print(json.dumps({'probe_copy_traffic_MB': mb, 'ctas_per_sm': cps, 'us': round(us, 2),
                          'GBps': round(mb * 1e3 / us, 1)}))
    del srcs, dsts
# launch floor: a 1-element kernel
a = torch.zeros(1, device='cuda')
us = timed([lambda: a.add_(1.0)] * 8, reps=50)
print(json.dumps({'empty_kernel_us_in_graph_chain': round(us, 2)}))
