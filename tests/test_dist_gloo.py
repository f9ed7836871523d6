"""CPU suite: the N>1 data path on two gloo ranks (127.0.0.1).

The batch is sharded along B, each rank evaluates its shard and the packed loss scalars are all-reduced once
(di_engine_b200.parallel).  On the GPUs the per-shard evaluation is the CUDA path; here -- no GPU in this container --
the per-shard evaluator is the CPU oracle, which is exactly what lets the test assert the sharding/collective logic:
mean-of-equal-shard-means == full-batch loss, gae shards concatenate to the full-batch advantage bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

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
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    import di_engine_b200 as b2
    from di_engine_b200 import parallel
    from oracle import rl_oracle
    import bench

    T, B, N = 32, 64, 6
    full = bench.make_batch(7, T=T, B=B, N=N)
    tm = {'value', 'next_value', 'reward', 'done', 'traj_flag'}
    # the PPO tensors are flattened (T*B, ...): view them time-major to shard along B like the trajectory tensors
    tb = {k: (v.view(T, B, *v.shape[1:]) if k not in tm else v) for k, v in full.items()}
    shard = parallel.shard_trajectory_batch(tb, rank, world, set(tb.keys()), set())
    assert shard['value'].shape == (T, B // world)

    adv = rl_oracle.gae(shard['value'], shard['next_value'].clone(), shard['reward'], shard['done'],
                        shard['traj_flag'], bench.GAMMA, bench.LAMBDA)
    flat = lambda x: x.reshape(-1, *x.shape[2:])  # noqa: E731
    out = rl_oracle.ppo_error(flat(shard['logit_new']), flat(shard['logit_old']), flat(shard['action']),
                              flat(shard['value_new']), flat(shard['value_old']), adv.reshape(-1),
                              flat(shard['return_']), None, None, bench.CLIP, True, None)
    red = parallel.LossAllReduce(6, 'cpu')
    work = red.reduce(list(out[:4]) + [out[4], out[5]], async_op=True)  # overlappable handle, as bench.py uses it
    reduced = red.finish(work).clone()
    # every rank ends with identical reduced values
    gathered = [torch.zeros_like(reduced) for _ in range(world)]
    dist.all_gather(gathered, reduced)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    advs = [torch.zeros_like(adv) for _ in range(world)]
    dist.all_gather(advs, adv)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r0.npz'), reduced=reduced.numpy(), adv=torch.cat(advs, 1).numpy())
    with pytest.raises(ValueError):
        parallel.shard_bounds(B + 1, rank, world)
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_step_matches_full_batch(tmp_path):
    import bench
    from oracle import rl_oracle
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    z = np.load(tmp_path / 'r0.npz')
    T, B, N = 32, 64, 6
    full = bench.make_batch(7, T=T, B=B, N=N)
    adv = rl_oracle.gae(full['value'], full['next_value'].clone(), full['reward'], full['done'], full['traj_flag'],
                        bench.GAMMA, bench.LAMBDA)
    assert np.array_equal(z['adv'], adv.numpy())  # column shards are independent: bit-identical advantage
    out = rl_oracle.ppo_error(full['logit_new'], full['logit_old'], full['action'], full['value_new'],
                              full['value_old'], adv.reshape(-1), full['return_'], None, None, bench.CLIP, True, None)
    want = np.array([float(x) for x in out[:4]] + [out[4], out[5]], dtype=np.float32)
    assert np.allclose(z['reduced'], want, rtol=1e-5, atol=1e-6), (z['reduced'], want)


def test_reference_arm_runs_only_on_rank0(tmp_path, monkeypatch, capsys):
    """bench.py --impl reference under torchrun: rank 0 prints the line, other ranks exit without work."""
    import bench
    import argparse
    args = argparse.Namespace(gpus=2, steps=1, warmup=1, impl='reference', unfused=False)
    monkeypatch.setenv('RANK', '1')
    bench.run_reference(args)
    assert capsys.readouterr().out == ''
