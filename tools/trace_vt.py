"""Timeline of the one-launch V-trace kernel (csrc/vtws.cu) at config E: per-CTA globaltimer stamps -> medians (us)."""
import os, sys
os.environ['B200RL_FUSED_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
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
rows = []
for j in range(8):
    rows += [(48 + 2 * j, 'load: stage free   %d' % j), (49 + 2 * j, 'load: issued       %d' % j),
             (4 * j, 'cons: A landed     %d' % j), (4 * j + 1, 'cons: A done       %d' % j),
             (32 + 2 * j, 'scan: IS ready     %d' % j), (33 + 2 * j, 'scan: vs published %d' % j),
             (4 * j + 2, 'cons: B vs ready   %d' % j), (4 * j + 3, 'cons: B done       %d' % j)]
for k, n in rows:
    c = tr[:, k]
    c = c[~np.isnan(c)]
    if len(c):
        print('%-24s min %7.2f  median %7.2f  max %7.2f  (n=%d)' % (n, c.min(), np.median(c), c.max(), len(c)))
