"""Where does the time of the small TD kernels go?  Variants of the config-B / config-C forward launch, CUDA-graph replay."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def timed(fns, reps=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


res = {}
for B in (64, 512, 4096):
    wl = bench.WorkloadC(B=B)
    sets = [wl.device_step(wl.make_batch(i), 'cuda:0') for i in range(4)]
    res['dntd_fwd_grad_B%d' % B] = round(timed([s.fwd_grad for s in sets]), 2)
    grads = [s.grad for s in sets]
    for s in sets:
        s.grad = None
    res['dntd_fwd_nograd_B%d' % B] = round(timed([s.fwd_grad for s in sets]), 2)
    for s, g in zip(sets, grads):
        s.grad = g
    res['dntd_bwd_check_B%d' % B] = round(timed([s.bwd_check for s in sets]), 2)
for B in (64, 512, 4096):
    wl = bench.WorkloadB(B=B)
    sets = [wl.device_step(wl.make_batch(i), 'cuda:0') for i in range(4)]
    res['qntd_fwd_grad_B%d' % B] = round(timed([s.fwd_grad for s in sets]), 2)
    res['qntd_bwd_check_B%d' % B] = round(timed([s.bwd_check for s in sets]), 2)
print(json.dumps(res))
