"""Timeline of the default column-tile kernel (csrc/colws.cu; B200RL_GAE_PPO_IMPL picks others) at config D: per-CTA globaltimer stamps -> medians (us)."""
import os, sys
os.environ['B200RL_FUSED_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
sets = [bench.DeviceStep(bench.make_batch(i), 'cuda:0', fused='onepass') for i in range(3)]
for _ in range(10):
    for s in sets:
        s.gae_ppo_fwd_grad()
torch.cuda.synchronize()
ws = sets[0].ws.view(torch.int64)
tr = ws[65536 // 2: 65536 // 2 + 32 * 512].cpu().numpy().reshape(-1, 32)
grid = int((tr[:, 0] != 0).sum())
tr = tr[:grid].astype(np.float64)
t0 = tr[:, 0].min()
tr = np.where(tr > 0, (tr - t0) / 1e3, np.nan)
names = {0: 'start', 19: 'scan: chunk0 published'}
for j in range(6):
    names[1 + 3 * j] = 'cons: chunk%d landed' % j
    names[2 + 3 * j] = 'cons: chunk%d adv ready' % j
    names[3 + 3 * j] = 'cons: chunk%d computed' % j
    names[20 + 2 * j] = 'prod: chunk%d done seen' % j
    names[21 + 2 * j] = 'prod: chunk%d refilled' % j
print('grid', grid)
order = [0, 19] + [k for j in range(6) for k in (1 + 3 * j, 2 + 3 * j, 3 + 3 * j, 20 + 2 * j, 21 + 2 * j)]
for k in order:
    c = tr[:, k]
    c = c[~np.isnan(c)]
    if len(c):
        print('%-26s min %7.2f  median %7.2f  max %7.2f  (n=%d)' % (names[k], c.min(), np.median(c), c.max(), len(c)))
