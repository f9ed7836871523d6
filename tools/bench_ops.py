"""Per-operator micro-benchmarks at the BASELINE.json configs (device-resident inputs, CUDA-graph replay over rotated
buffer sets larger than L2, CUDA events).  Prints one JSON line per op: us per call and algorithmic GB/s.

    python tools/bench_ops.py [gae] [qntd] [dntd] [vtrace] [tdl] [upgo] [gae1d] [quantile] [retrace] [happo]
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
        elif op in ('qrdqn', 'iqn', 'fqf'):
            data = getattr(r, op + '_nstep_td_data')(*[td[k] for k in cases.QUANTILE_FIELDS[op]])
            loss = getattr(r, op + '_nstep_td_error')(data, value_gamma=td.get('value_gamma'), **p)[0]
        elif op == 'retrace':
            r.compute_q_retraces(*td.values(), **p)
            return
        elif op == 'happo':
            l, _ = r.happo_error(r.happo_data(*[td[k] for k in cases.HAPPO_FIELDS]), **p)
            torch.autograd.backward([l.policy_loss, l.value_loss, l.entropy_loss], [G1, G05, GM001])
            return
        elif op == 'gae1d':
            r.gae_returns(r.gae_data(td['value'], td['next_value'], td['reward'], td['done'], td['traj_flag']), 0.99, 0.95, 1.7)
            return
        loss.backward()

    us = timed([lambda td=td, p=p: call(td, p) for td, p in sets], reps=200)
    report(name + ' fwd+bwd (public API, graph-captured kernels)', us, alg_bytes, units, unit_name)


if __name__ == '__main__':
    which = sys.argv[1:] or ['gae', 'qntd', 'dntd', 'vtrace', 'tdl', 'upgo', 'gae1d', 'quantile', 'retrace', 'happo']
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
    if 'gae1d' in which:
        cases.GRAD_INPUTS['gae1d'] = []
        bench_api('gae_returns 1-D T=3200 (the PPO learner call: value-norm, returns, both statistics)', 'gae1d',
                  lambda i: ('gae1d', ) + cases.gae_case(i, 3200, 1, one_d=True, p_done=0.0025)[1:], 36 * 3200, 3200, 'transitions')
    if 'quantile' in which:
        bench_api('qrdqn_nstep_td_error B=64 N=6 num=200 n=3', 'qrdqn',
                  lambda i: cases.quantile_case(i, 'qrdqn', 64, 6, 200, 200, 3, weight='tensor'), 64 * (2 * 6 * 200 * 4 + 40), 64,
                  'samples')
        bench_api('iqn_nstep_td_error B=64 N=6 tau=32 n=3', 'iqn',
                  lambda i: cases.quantile_case(i, 'iqn', 64, 6, 32, 32, 3, weight='tensor'), 64 * (2 * 6 * 32 * 4 + 40), 64,
                  'samples')
    if 'retrace' in which:
        bench_api('compute_q_retraces T=64 B=8192 N=6', 'retrace', lambda i: cases.retrace_case(i, 64, 8192, 6),
                  64 * 8192 * (4 * 6 * 2 + 24), 64 * 8192, 'transitions')
    if 'happo' in which:
        bench_api('happo_error B=65536 N=6', 'happo', lambda i: cases.happo_case(i, 65536, 6, weight='tensor'), 108 * 65536, 65536,
                  'samples')
