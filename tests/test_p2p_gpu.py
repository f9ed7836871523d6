"""GPU suite, two ranks on two GPUs of one box (skipped when fewer are visible): the N>1 data plane.

  * ``b200rl_p2p_allreduce_mean`` (the separate one-CTA exchange kernel) over NVLink peer memory;
  * the exchange fused into the epilogue of the one-launch learner step (``b200rl_gae_ppo_fwd_grad_dp``): every rank runs
    its column shard, the mean of the rank means of the six loss scalars arrives in ``out_mean`` one step later (and after
    ``drain()`` for the last step) and equals the CPU oracle's full-batch losses; adv shards are bit-identical to the
    full-batch advantage; also under CUDA-graph replay.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'] = str(rank)
    os.environ['WORLD_SIZE'] = str(world)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import bench
    from di_engine_b200 import ops, parallel

    # ---- (1) the separate exchange kernel --------------------------------------------------------------------------
    red = parallel.P2PLossAllReduce(6, dev)
    for it in range(50):
        src = torch.arange(8, device=dev, dtype=torch.float32) * (rank + 1) + it
        got = red.reduce(src).clone().cpu()
        want = torch.arange(6, dtype=torch.float32) * (sum(r + 1 for r in range(world)) / world) + it
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-6), (rank, it, got, want)

    # ---- (2) the exchange fused into the learner step ---------------------------------------------------------------
    T, B, N = 64, 256, 6
    steps = 5
    x = parallel.FusedLossExchange(dev)
    results = []
    wl = bench.WorkloadD(B=B // world, T=T, N=N)
    devsteps = []
    for s in range(steps):
        full = bench.make_batch(100 + s, T=T, B=B, N=N)
        tm = {'value', 'next_value', 'reward', 'done', 'traj_flag'}
        tb = {k: (v.view(T, B, *v.shape[1:]) if k not in tm else v) for k, v in full.items()}
        shard = parallel.shard_trajectory_batch(tb, rank, world, set(tb.keys()), set())
        flat = {k: (v.reshape(-1, *v.shape[2:]) if k not in tm else v).contiguous() for k, v in shard.items()}
        devsteps.append(bench.DeviceStepD(wl, flat, str(dev), x))
    for s in range(steps):
        devsteps[s]()
        torch.cuda.synchronize()
        if s > 1:  # step s consumed the mean of step s-2 (step s-1's was published by step s's kernel)
            results.append(x.out_mean[:6].clone().cpu())
    x.drain()
    torch.cuda.synchronize()
    results.append(x.out_mean[8:14].clone().cpu())  # step Q-1
    results.append(x.out_mean[:6].clone().cpu())    # step Q
    advs = [torch.zeros_like(devsteps[0].adv) for _ in range(world)]
    dist.all_gather(advs, devsteps[0].adv)
    local = torch.stack([d.out[:6].clone() for d in devsteps]).cpu()
    # graph replay: two more steps captured once, replayed three times; ranks stay in lock step through the mailboxes
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        devsteps[0]()
        devsteps[1]()
        x.drain()
        stream.synchronize()
        dist.barrier()
        with torch.cuda.graph(g, stream=stream):
            devsteps[0]()
            devsteps[1]()
            x.drain()
        for _ in range(3):
            g.replay()
    stream.synchronize()
    replayed = x.out_mean[:6].clone().cpu()
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r0.npz'), means=torch.stack(results).numpy(), adv=torch.cat(advs, 1).cpu().numpy(),
                 local=local.numpy(), replayed=replayed.numpy())
    gathered = [torch.zeros(steps, 6, device=dev) for _ in range(world)]
    dist.all_gather(gathered, torch.stack(results).to(dev))
    assert all(torch.equal(t, gathered[0]) for t in gathered), 'every rank must end with identical means'
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason='needs two GPUs')
def test_two_rank_exchange_kernels(tmp_path):
    import bench
    from oracle import rl_oracle
    port = _free_port()
    ctx = mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=False)
    import time
    t0 = time.time()
    while not ctx.join(timeout=5):
        assert time.time() - t0 < 240, 'two-rank worker timed out'
    for p in ctx.processes:
        assert p.exitcode == 0, p.exitcode
    z = np.load(tmp_path / 'r0.npz')
    T, B, N = 64, 256, 6
    for s in range(5):
        full = bench.make_batch(100 + s, T=T, B=B, N=N)
        adv = rl_oracle.gae(full['value'], full['next_value'].clone(), full['reward'], full['done'], full['traj_flag'],
                            bench.GAMMA, bench.LAMBDA)
        if s == 0:
            assert np.array_equal(z['adv'], adv.numpy())
        out = rl_oracle.ppo_error(full['logit_new'], full['logit_old'], full['action'], full['value_new'],
                                  full['value_old'], adv.reshape(-1), full['return_'], None, None, bench.CLIP, True, None)
        want = np.array([float(v) for v in out[:4]] + [out[4], out[5]], dtype=np.float32)
        assert np.allclose(z['means'][s], want, rtol=1e-5, atol=1e-6), (s, z['means'][s], want)
    # after the replays (each ends with a drain) the latest mean is that of the graph's last step, buffer set 1
    full = bench.make_batch(101, T=T, B=B, N=N)
    adv = rl_oracle.gae(full['value'], full['next_value'].clone(), full['reward'], full['done'], full['traj_flag'],
                        bench.GAMMA, bench.LAMBDA)
    out = rl_oracle.ppo_error(full['logit_new'], full['logit_old'], full['action'], full['value_new'], full['value_old'],
                              adv.reshape(-1), full['return_'], None, None, bench.CLIP, True, None)
    want = np.array([float(v) for v in out[:4]] + [out[4], out[5]], dtype=np.float32)
    assert np.allclose(z['replayed'], want, rtol=1e-5, atol=1e-6)
