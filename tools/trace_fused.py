"""Timeline of the one-pass kernel (B200RL_FUSED_TRACE=1): per-CTA timestamps -> summary."""
import os, sys
os.environ['B200RL_FUSED_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
s = bench.DeviceStep(bench.make_batch(0), 'cuda:0', fused='onepass')
for _ in range(20):
    s.gae_ppo_fwd_grad()
torch.cuda.synchronize()
ws = s.ws.view(torch.int64)  # float32 words -> int64 pairs
tr = ws[65536 // 2: 65536 // 2 + 8 * 1024].cpu().numpy().reshape(-1, 8)
grid = int((tr[:, 0] != 0).sum())
tr = tr[:grid].astype(np.float64)
t0 = tr[:, 0].min()
tr = (tr - t0) / 1e3
names = ['start', 'gae_chunk0', 'gae_chunk1', 'gae_chunk2', 'gae_chunk3', 'cta_end', 'first_tile_ready', 'ppo_done']
print('grid', grid)
for k, n in enumerate(names):
    c = tr[:256, k] if k in (1, 2, 3, 4) else tr[:, k]
    if k == 5:
        c = c[c > 0]
    print('%-18s min %7.2f  median %7.2f  max %7.2f us  (n=%d)' % (n, c.min(), np.median(c), c.max(), len(c)))
print('non-GAE CTAs first tile ready: median %.2f' % np.median(tr[256:, 6]))
print('GAE CTAs     first tile ready: median %.2f' % np.median(tr[:256, 6]))
print('non-GAE CTAs ppo_done: median %.2f   GAE CTAs ppo_done: median %.2f' % (np.median(tr[256:, 7]), np.median(tr[:256, 7])))

