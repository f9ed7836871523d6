"""Per-operator micro-benchmarks at the BASELINE.json configs (device-resident inputs, CUDA-graph replay over rotated
buffer sets larger than L2, CUDA events).  Prints one JSON line per op: us per call and algorithmic GB/s.

    python tools/bench_ops.py [gae] [qntd] [dntd] [vtrace] [tdl] [upgo]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import di_engine_b200 as b2  # noqa: E402
from tests import cases  # noqa: E402

DEV = 'cuda:0'
G1 = torch.tensor(1.0, device=DEV)
G05 = torch.tensor(0.5, device=DEV)
GM001 = torch.tensor(-0.01, device=DEV)


def timed(fn_sets, reps=50):
    """fn_sets: list of zero-arg callables (one per buffer set). Returns us per call."""
    main = torch.cuda.Stream()
    with torch.cuda.stream(main):
        for f in fn_sets:
            f()
            f()
        main.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for f in fn_sets:
                f()
        # pre-heat ~0.2 s so that clocks and caches are in steady state, then measure >= `reps` replays and >= 50 ms
        import time
        t_end = time.perf_counter() + 0.2
        while time.perf_counter() < t_end:
            for _ in range(20):
                g.replay()
            main.synchronize()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(reps, 400)
            main.synchronize()
            e0.record(main)
            for _ in range(n):
                g.replay()
            e1.record(main)
            main.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (n * len(fn_sets))
            best = us if best is None else min(best, us)
    return best


def report(name, us, alg_bytes, units, unit_name):
    print(json.dumps({'op': name, 'us_per_call': round(us, 3), 'alg_GBps': round(alg_bytes / us / 1e3, 1),
                      'units_per_s': units / (us * 1e-6), 'unit': unit_name}))


def bench_gae(T=128, B=4096, nsets=12):
    from di_engine_b200 import ops
    sets = []
    for i in range(nsets):
        _, t, p = cases.gae_case(i, T, B, p_done=0.01)
        t = {k: v.to(DEV) for k, v in t.items()}
        sets.append(lambda t=t: ops.gae_(t['value'], t['next_value'], t['reward'], t['done'], t['traj_flag'], 0.99,
                                         0.95, 1))
    report('gae T=%d B=%d' % (T, B), timed(sets), 24 * T * B, T * B, 'transitions')


def bench_api(name, op, mk, alg_bytes, units, unit_name, nsets=4):
    """forward + backward through the PUBLIC API on device-resident tensors, captured once into a CUDA graph (so the
    figure is the kernels' time, not the Python/ctypes dispatch), replayed over rotated input sets."""
    import di_engine_b200.rl_utils.ppo as _p
    _p.LAZY_INFO = True
    sets = []
    for i in range(nsets):
        _, t, p = mk(i)
        td = cases.prepare(op, t, DEV)
        p = dict(p)
        if 'gamma' in p and isinstance(p['gamma'], list):
            p['gamma'] = [g.to(DEV) for g in p['gamma']]
        sets.append((td, p))

    def call(td, p):
        for k in cases.GRAD_INPUTS[op]:
            td[k].grad = None
        r = b2.rl_utils
        if op == 'qntd':
            data = r.q_nstep_td_data(*[td[k] for k in ('q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done',
                                                        'weight')])
            pp = dict(p); g = pp.pop('gamma')
            loss = r.q_nstep_td_error(data, g, value_gamma=td.get('value_gamma'), **pp)[0]
        elif op == 'dntd':
            data = r.dist_nstep_td_data(td['dist'], td['next_n_dist'], td['act'], td['next_n_act'], td['reward'],
                                        td['done'], td['weight'])
            loss = r.dist_nstep_td_error(data, value_gamma=td.get('value_gamma'), **p)[0]
        elif op == 'vtrace':
            data = r.vtrace_data(td['target_output'], td['behaviour_output'], td['action'], td['value'], td['reward'],
                                 td['weight'])
            l = r.vtrace_error_discrete_action(data, **p)
            # upstream gradients as ready-made device scalars: no torch arithmetic kernels in the measured graph
            torch.autograd.backward([l.policy_loss, l.value_loss, l.entropy_loss], [G1, G05, GM001])
            return
        elif op == 'td_lambda':
            loss = r.td_lambda_error(r.td_lambda_data(td['value'], td['reward'], td['weight']), **p)
        elif op == 'upgo':
            loss = r.upgo_loss(td['target_output'], td['rhos'], td['action'], td['rewards'], td['bootstrap_values'],
                               td['mask'])
        loss.backward()

    us = timed([lambda td=td, p=p: call(td, p) for td, p in sets], reps=200)
    report(name + ' fwd+bwd (public API, graph-captured kernels)', us, alg_bytes, units, unit_name)


if __name__ == '__main__':
    which = sys.argv[1:] or ['gae', 'qntd', 'dntd', 'vtrace', 'tdl', 'upgo']
    b2.rl_utils.td.CHECK_DIST_POSITIVE = False
    if 'gae' in which:
        bench_gae()
        bench_gae(T=1024, B=64, nsets=8)
    if 'qntd' in which:
        bench_api('q_nstep_td_error B=512 N=6 n=3', 'qntd',
                  lambda i: cases.qntd_case(i, 512, 6, 3, value_gamma='tensor', gamma=0.99, done='bern'), 120 * 512,
                  512, 'samples')
    if 'dntd' in which:
        bench_api('dist_nstep_td_error B=512 N=6 atoms=51 n=3', 'dntd',
                  lambda i: cases.dntd_case(i, 512, 6, 51, 3, gamma=0.99, value_gamma='tensor'), 1670 * 512, 512,
                  'samples')
    if 'vtrace' in which:
        bench_api('vtrace T=64 B=8192 N=6', 'vtrace',
                  lambda i: cases.vtrace_case(i, 64, 8192, 6, gamma=0.99, lambda_=0.95), 96 * 64 * 8192, 64 * 8192,
                  'transitions')
    if 'tdl' in which:
        bench_api('td_lambda T=1024 B=64', 'td_lambda', lambda i: cases.td_lambda_case(i, 1024, 64), 16 * 1024 * 64,
                  1024 * 64, 'transitions')
    if 'upgo' in which:
        bench_api('upgo T=256 B=256 N=256', 'upgo', lambda i: cases.upgo_case(i, 256, 256, 256),
                  (8 * 256 + 20) * 65536, 65536, 'transitions')
