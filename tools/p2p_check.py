"""Multi-GPU check of the NVLink peer-memory all-reduce (run under torchrun, one rank per GPU):

    torchrun --nproc-per-node 2 tools/p2p_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():

    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from di_engine_b200.parallel import P2PLossAllReduce, LossAllReduce

    red = P2PLossAllReduce(6, dev)
    ok = True
    for it in range(200):
        src = torch.arange(8, device=dev, dtype=torch.float32) * (rank + 1) + it
        out = red.reduce(src).clone()
        want = torch.arange(6, dtype=torch.float32) * (sum(r + 1 for r in range(world)) / world) + it
        if not torch.allclose(out.cpu(), want, rtol=1e-6, atol=1e-6):
            ok = False
            print('rank', rank, 'mismatch at', it, out.cpu(), want)
            break
    # graph capture + timing
    src = torch.ones(8, device=dev) * (rank + 1)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        red.reduce(src)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            red.reduce(src)
        for _ in range(20):
            g.replay()
        s.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(500):
            g.replay()
        e1.record(s)
        s.synchronize()
    p2p_us = e0.elapsed_time(e1) * 1e3 / 500
    assert abs(red.buf[0].item() - sum(r + 1 for r in range(world)) / world) < 1e-6
    nc = LossAllReduce(6, dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(200):
        nc.reduce([1.0] * 6)
    torch.cuda.synchronize()
    nccl_us = (time.perf_counter() - t0) * 1e6 / 200
    if rank == 0:
        print('p2p allreduce ok=%s  world=%d  graph-replayed p2p: %.2f us/op   eager NCCL LossAllReduce: %.1f us/op' %
              (ok, world, p2p_us, nccl_us), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
