import os, sys
os.environ['B200RL_FUSED_TRACE'] = '1'
sys.path.insert(0, '/root/repo')
import numpy as np, torch
sys.argv = sys.argv[:1]
import tools.exp_vt as ev
ev.main()
torch.cuda.synchronize()
for x in ev.sets[:2]:
    x.fused()
torch.cuda.synchronize()
ws = ev.ws.view(torch.int64)
tr = ws[65536 // 2: 65536 // 2 + 64 * 512].cpu().numpy().reshape(-1, 64)
grid = int((tr[:, 0] != 0).sum())
tr = tr[:grid].astype(np.float64)
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, (tr - t0) / 1e3, np.nan)
print('grid', grid)
def row(k, n):
    c = tr[:, k]; c = c[~np.isnan(c)]
    if len(c): print('%-28s min %6.2f med %6.2f max %6.2f' % (n, c.min(), np.median(c), c.max()))
row(0, 'A landed 0 (warp0)'); row(1, 'A done 0 (warp0)')
for w in range(8): row(24 + w, 'A(0) done warp %d' % w)
a = tr[:, 24:32]; print('per-CTA spread of A(0) done across warps: median %.2f max %.2f' % (np.nanmedian(np.nanmax(a,1)-np.nanmin(a,1)), np.nanmax(np.nanmax(a,1)-np.nanmin(a,1))))
print('per-CTA last warp A(0) done: median %.2f' % np.nanmedian(np.nanmax(a,1)))
row(32, 'scan IS ready 0'); row(33, 'scan vs published 0'); row(2, 'B vs ready 0 (warp0)')
for w in range(8): row(56 + w, 'B(0) done warp %d' % w)
row(48, 'load stage free 0'); row(49, 'load issued 0'); row(51, 'load issued 1'); row(53, 'load issued 2'); row(55, 'load issued 3')
