"""Column-tile one-launch learner step (csrc/coltile.cu) against the other step variants at config D: a correctness
cross-check (advantages bit-identical, losses / gradients against the three-kernel path) and then timings.

    B200RL_COL_TC=16 python tools/exp_col.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from di_engine_b200 import ops
from tools.bench_ops import timed

L = ops.lib()
res = {'tc': os.environ.get('B200RL_COL_TC', '16'), 'knobs': {k: v for k, v in os.environ.items() if k.startswith('B200RL_')}}
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused=True) for i in range(6)]


def snapshot(s):
    torch.cuda.synchronize()
    return dict(adv=s.adv.clone(), out=s.out.clone(), gl=s.grad_logit.clone(), gv=s.grad_value.clone(),
                nv=s.b['next_value'].clone())


def reset(s):
    s.b['next_value'].copy_(s.nv0)
    s.adv.zero_(); s.out.zero_(); s.grad_logit.zero_(); s.grad_value.zero_()


s = sets[0]
reset(s); s.gae(); s.ppo_fwd_grad(); ref = snapshot(s)
for name, impl in (('row', 1), ('col', 2), ('colcp', 3), ('coltma', 4)):
    L.b200rl_gae_ppo_set_impl(impl)
    reset(s); s.gae_ppo_fwd_grad(); got = snapshot(s)
    ok = {
        'adv': bool(torch.equal(got['adv'], ref['adv'])),
        'nv': bool(torch.equal(got['nv'], ref['nv'])),
        'out': bool(torch.allclose(got['out'][:6], ref['out'][:6], rtol=1e-5, atol=1e-6)),
        'gl_maxdiff': float((got['gl'] - ref['gl']).abs().max()),
        'gv_maxdiff': float((got['gv'] - ref['gv']).abs().max()),
        'gl_absmax': float(ref['gl'].abs().max()),
    }
    res['check_' + name] = ok
for s in sets:
    reset(s)
torch.cuda.synchronize()

res['gae_us'] = round(timed([s.gae for s in sets], reps=30), 2)
res['ppo_fwd_grad_us'] = round(timed([s.ppo_fwd_grad for s in sets], reps=30), 2)


def three(s):
    s.gae(); s.ppo_fwd_grad(); s.ppo_bwd_check()


def one(s):
    s.gae_ppo_fwd_grad(); s.ppo_bwd_check()


res['step3_us'] = round(timed([lambda s=s: three(s) for s in sets], reps=30), 2)
for name, impl in (('row', 1), ('col', 2), ('colcp', 3), ('coltma', 4)):
    L.b200rl_gae_ppo_set_impl(impl)
    res['onepass_%s_us' % name] = round(timed([s.gae_ppo_fwd_grad for s in sets], reps=30), 2)
    res['step1_%s_us' % name] = round(timed([lambda s=s: one(s) for s in sets], reps=30), 2)
L.b200rl_gae_ppo_set_impl(0)
res['step1_col_frac_of_peak'] = round(67108864 / (res['step1_col_us'] * 1e-6) / 1e9 / bench.load_peaks()[0], 4)
print(json.dumps(res))
