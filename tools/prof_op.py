"""Eager forward + backward of one operator case through the public API, for ncu launch lists:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/x.csv python tools/prof_op.py upgo
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import di_engine_b200 as b2  # noqa: E402
from tests import cases  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'upgo'
mk = {
    'upgo': lambda i: cases.upgo_case(i, 256, 256, 256),
    'td_lambda': lambda i: cases.td_lambda_case(i, 1024, 64),
    'qrdqn': lambda i: cases.quantile_case(i, 'qrdqn', 64, 6, 200, 200, 3, weight='tensor'),
    'retrace': lambda i: cases.retrace_case(i, 64, 8192, 6),
}[which]
for i in range(3):
    op, t, p = mk(i)
    res = cases.run_api(b2.rl_utils, op, t, p, device='cuda:0')
torch.cuda.synchronize()
print('ran', which)
