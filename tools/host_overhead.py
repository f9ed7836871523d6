"""Host-side cost of the public API at the learner's REAL call sizes (PPO minibatches of 64 / 320 rows,
dizoo/atari/config/serial/pong/pong_ppo_config.py:29; one DQN batch of 64; the 1-D gae over n_sample = 3200):
wall-clock microseconds per call, forward + backward, device synchronised once per timed loop -- so the figure is
max(host issue time, device time) per call, which at these sizes is the host.  Next to it: the reference's own torch functions
on the same CUDA tensors (oracle/_ref archive) and on the host CPU.

    python tools/host_overhead.py > profiles/rNN_host_overhead.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import di_engine_b200 as b2  # noqa: E402
from oracle import ref_loader  # noqa: E402
from tests import cases  # noqa: E402

DEV = 'cuda:0'
ref = ref_loader.load() if ref_loader.available() else None


def wall(fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e6 / n


def ppo_call(api, t):
    ln = t['logit_new'].detach().requires_grad_(True)
    vn = t['value_new'].detach().requires_grad_(True)
    loss, info = api.ppo_error(api.ppo_data(ln, t['logit_old'], t['action'], vn, t['value_old'], t['adv'], t['return_'], None,
                                            None), 0.2)
    (loss.policy_loss + 0.5 * loss.value_loss - 0.01 * loss.entropy_loss).backward()


def qntd_call(api, t):
    q = t['q'].detach().requires_grad_(True)
    loss, per = api.q_nstep_td_error(api.q_nstep_td_data(q, t['next_n_q'], t['action'], t['next_n_action'], t['reward'],
                                                         t['done'], None), 0.99, nstep=3)
    loss.backward()


def gae_call(api, t):
    api.gae(api.gae_data(t['value'], t['next_value'].clone(), t['reward'], t['done'], t['traj_flag']), 0.99, 0.95)


res = {}
for rows in (64, 320):
    _, t, _ = cases.ppo_case(1, rows, 6)
    td = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in t.items()}
    res['ppo_error_fwd_bwd_%d_rows_b200_us' % rows] = round(wall(lambda: ppo_call(b2, td)), 1)
    b2.rl_utils.ppo.LAZY_INFO = True
    res['ppo_error_fwd_bwd_%d_rows_b200_lazy_info_us' % rows] = round(wall(lambda: ppo_call(b2, td)), 1)
    b2.rl_utils.ppo.LAZY_INFO = False
    if ref is not None:
        res['ppo_error_fwd_bwd_%d_rows_reference_cuda_us' % rows] = round(wall(lambda: ppo_call(ref, td), 100), 1)
        torch.set_num_threads(1)
        res['ppo_error_fwd_bwd_%d_rows_reference_cpu_us' % rows] = round(wall(lambda: ppo_call(ref, t), 100), 1)
_, t, _ = cases.qntd_case(2, 64, 6, 3, done='bern')
td = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in t.items()}
res['q_nstep_td_error_fwd_bwd_64_b200_us'] = round(wall(lambda: qntd_call(b2, td)), 1)
if ref is not None:
    res['q_nstep_td_error_fwd_bwd_64_reference_cuda_us'] = round(wall(lambda: qntd_call(ref, td), 100), 1)
    res['q_nstep_td_error_fwd_bwd_64_reference_cpu_us'] = round(wall(lambda: qntd_call(ref, t), 100), 1)
g = torch.Generator().manual_seed(3)
T = 3200
t = dict(value=torch.randn(T, generator=g), next_value=torch.randn(T, generator=g), reward=torch.randn(T, generator=g),
         done=(torch.rand(T, generator=g) < 0.003).float())
t['traj_flag'] = t['done'].clone()
t['traj_flag'][399::400] = 1.0
td = {k: v.to(DEV) for k, v in t.items()}
res['gae_1d_T3200_b200_us'] = round(wall(lambda: gae_call(b2, td)), 1)
if ref is not None:
    res['gae_1d_T3200_reference_cuda_us'] = round(wall(lambda: gae_call(ref, td), 5), 1)
    res['gae_1d_T3200_reference_cpu_us'] = round(wall(lambda: gae_call(ref, t), 5), 1)
# device time of the 1-D gae alone (graph replay)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gae_call(b2, td)
    s.synchronize()
    gr = torch.cuda.CUDAGraph()
    nv = td['next_value'].clone()
    with torch.cuda.graph(gr, stream=s):
        b2.gae(b2.gae_data(td['value'], nv, td['reward'], td['done'], td['traj_flag']), 0.99, 0.95)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        gr.replay()
    e0.record(s)
    for _ in range(200):
        gr.replay()
    e1.record(s)
    s.synchronize()
res['gae_1d_T3200_b200_device_us'] = round(e0.elapsed_time(e1) * 1e3 / 200, 2)
print(json.dumps(res, indent=1))
