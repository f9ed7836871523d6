"""V-trace at config E (T=64, B=8192, N=6) through the C ABI: the one-launch kernel of csrc/vtws.cu (+ its verification
launch) against the rows / scan / backward kernels of csrc/pg.cu.  Device-resident inputs, CUDA-graph replay over rotated
buffer sets (6 x 50 MB > L2), CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from di_engine_b200 import ops
from tests import cases
from tools.bench_ops import timed

DEV = 'cuda:0'
T, B, N = 64, 8192, 6
L = ops.lib()
ws = ops.workspace(torch.device(DEV))


class VtStep:
    def __init__(self, seed):
        _, t, _ = cases.vtrace_case(seed, T, B, N)
        self.t = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in t.items()}
        self.out = torch.zeros(4, device=DEV)
        self.gl = torch.empty(T, B, N, device=DEV)
        self.gv = torch.empty(T + 1, B, device=DEV)
        self.g_used = torch.zeros(3, device=DEV)
        self.hint = torch.tensor([1.0, 0.5, -0.01], device=DEV)
        self.g = [torch.tensor(x, device=DEV) for x in (1.0, 0.5, -0.01)]
        self.lp, self.cpg, self.dv = (torch.empty(T, B, device=DEV) for _ in range(3))

    def args(self):
        t = self.t
        return (ops.ptr(t['target_output']), ops.ptr(t['behaviour_output']), ops.ptr(t['action']), ops.ptr(t['value']),
                ops.ptr(t['reward']), None, T, B, N, 0.99, 0.95, 1.0, 1.0, 1.0)

    def fused(self):
        rc = L.b200rl_vtrace_fwd_grad(*self.args(), ops.ptr(self.hint), 0, None, None, None, ops.ptr(self.g_used), None,
                                      ops.ptr(self.out), ops.ptr(self.gl), ops.ptr(self.gv), ops.ptr(ws), ws.numel() * 4,
                                      ops.stream_ptr())
        assert rc == 0, rc

    def verify(self):
        rc = L.b200rl_vtrace_fwd_grad(*self.args(), None, 1, ops.ptr(self.g[0]), ops.ptr(self.g[1]), ops.ptr(self.g[2]),
                                      ops.ptr(self.g_used), ops.ptr(self.hint), None, ops.ptr(self.gl), ops.ptr(self.gv),
                                      ops.ptr(ws), ws.numel() * 4, ops.stream_ptr())
        assert rc == 0, rc

    def legacy_fwd(self):
        rc = L.b200rl_vtrace_fwd(*self.args(), ops.ptr(self.out), ops.ptr(self.lp), ops.ptr(self.cpg), ops.ptr(self.dv),
                                 ops.ptr(ws), ws.numel() * 4, ops.stream_ptr())
        assert rc == 0, rc

    def legacy_bwd(self):
        t = self.t
        rc = L.b200rl_vtrace_bwd(ops.ptr(t['target_output']), ops.ptr(t['action']), None, ops.ptr(self.cpg),
                                 ops.ptr(self.dv), ops.ptr(self.g[0]), ops.ptr(self.g[1]), ops.ptr(self.g[2]), T, B, N,
                                 ops.ptr(self.gl), ops.ptr(self.gv), ops.stream_ptr())
        assert rc == 0, rc


sets = [VtStep(i) for i in range(6)]
def main():
    s = sets[0]
    s.legacy_fwd(); s.legacy_bwd(); torch.cuda.synchronize()
    ref = (s.out.clone(), s.gl.clone(), s.gv.clone())
    s.gl.zero_(); s.gv.zero_(); s.out.zero_()
    s.fused(); s.verify(); torch.cuda.synchronize()
    res = {'check': {'out': bool(torch.allclose(s.out[:3], ref[0][:3], rtol=1e-5, atol=1e-6)),
                     'gl_maxdiff': float((s.gl - ref[1]).abs().max()), 'gl_absmax': float(ref[1].abs().max()),
                     'gv_maxdiff': float((s.gv - ref[2]).abs().max())}}


    def both(x):
        x.fused(); x.verify()


    def legacy(x):
        x.legacy_fwd(); x.legacy_bwd()


    res['fused_us'] = round(timed([x.fused for x in sets], reps=30), 2)
    res['fused_plus_verify_us'] = round(timed([lambda x=x: both(x) for x in sets], reps=30), 2)
    res['legacy_fwd_bwd_us'] = round(timed([lambda x=x: legacy(x) for x in sets], reps=30), 2)
    peak = bench.load_peaks()[0]
    res['fused_plus_verify_frac'] = round(96 * T * B / (res['fused_plus_verify_us'] * 1e-6) / 1e9 / peak, 4)
    res['legacy_frac'] = round(96 * T * B / (res['legacy_fwd_bwd_us'] * 1e-6) / 1e9 / peak, 4)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
