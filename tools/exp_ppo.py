"""Tuning experiments on the PPO tile kernel (env knobs B200RL_PPO_DBG / B200RL_PPO_CTAS): time fwd_grad alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools.bench_ops import timed
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused=True) for i in range(6)]
for s in sets:
    s.gae()
torch.cuda.synchronize()
us = timed([s.ppo_fwd_grad for s in sets], reps=30)
us_f = timed([s.ppo_fwd for s in sets], reps=30)
print(json.dumps({'dbg': os.environ.get('B200RL_PPO_DBG', '0'), 'ctas': os.environ.get('B200RL_PPO_CTAS', '-'),
                  'fwd_grad_us': round(us, 2), 'fwd_us': round(us_f, 2)}))
