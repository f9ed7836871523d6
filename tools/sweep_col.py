"""Time the default one-launch step (csrc/colws.cu) under the B200RL_COL_* knobs of the environment: one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools.bench_ops import timed
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused='onepass') for i in range(6)]


def one(s):
    s.gae_ppo_fwd_grad(); s.ppo_bwd_check()


res = {'knobs': {k: v for k, v in os.environ.items() if k.startswith('B200RL_')}}
sets3 = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused=True) for i in range(6)]
res['control_step3_us'] = round(timed([s for s in sets3], reps=30), 2)
res['onepass_us'] = round(timed([s.gae_ppo_fwd_grad for s in sets], reps=30), 2)
res['step_us'] = round(timed([lambda s=s: one(s) for s in sets], reps=30), 2)
res['step_frac'] = round(67108864 / (res['step_us'] * 1e-6) / 1e9 / bench.load_peaks()[0], 4)
print(json.dumps(res))
