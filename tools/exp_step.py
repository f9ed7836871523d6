"""Tuning experiments: time the learner step variants under the env knobs (B200RL_PPO_RPT, B200RL_PPO_CTAS...)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools.bench_ops import timed
res = {'rpt': os.environ.get('B200RL_PPO_RPT', 'auto'), 'ctas': os.environ.get('B200RL_PPO_CTAS', '-')}
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused=True) for i in range(6)]
for s in sets:
    s.gae()
torch.cuda.synchronize()
res['ppo_fwd_grad_us'] = round(timed([s.ppo_fwd_grad for s in sets], reps=30), 2)
res['ppo_fwd_us'] = round(timed([s.ppo_fwd for s in sets], reps=30), 2)
res['gae_us'] = round(timed([s.gae for s in sets], reps=30), 2)
res['onepass_us'] = round(timed([s.gae_ppo_fwd_grad for s in sets], reps=30), 2)
def three(s):
    s.gae(); s.ppo_fwd_grad(); s.ppo_bwd_check()
def one(s):
    s.gae_ppo_fwd_grad(); s.ppo_bwd_check()
res['step3_us'] = round(timed([lambda s=s: three(s) for s in sets], reps=30), 2)
res['step1_us'] = round(timed([lambda s=s: one(s) for s in sets], reps=30), 2)
print(json.dumps(res))
