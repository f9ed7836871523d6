"""Randomised shapes (seeded) for the path's operators against the oracle: ragged tiles, odd action counts, tiny and prime sizes --
the places where tile kernels, fast paths and fallbacks hand over to each other.  gae / lambda-returns / Retrace stay bit-exact."""
import numpy as np
import pytest
import torch

import di_engine_b200 as b2
from oracle import rl_oracle
from tests import cases

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _draws(seed, n, *ranges):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        out.append(tuple(int(rng.choice(r)) if isinstance(r, (list, tuple)) else int(rng.randint(r.start, r.stop)) for r in ranges))
    return out


def _check(op, t, p, exact=False):
    want = cases.run_oracle(rl_oracle, op, t, p)
    got = cases.run_api(b2.rl_utils, op, t, p, device=DEV)
    if exact:
        cases.compare(got, want, exact=True)
    else:
        cases.compare(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('T,B,kind', _draws(1, 14, range(1, 200), range(1, 700), [0, 1, 2, 3]))
def test_fuzz_gae(T, B, kind):
    kw = [dict(), dict(done=None, traj=None), dict(done='bool', traj='bool', p_done=0.2), dict(traj=None, p_done=0.3)][kind]
    op, t, p = cases.gae_case(1000 + T + B, T, B, **kw)
    _check(op, t, p, exact=True)


@pytest.mark.parametrize('B,N,kind', _draws(2, 16, range(1, 5000), [2, 3, 5, 6, 7, 13, 18, 33, 64, 65, 130], [0, 1, 2, 3]))
def test_fuzz_ppo(B, N, kind):
    kw = [dict(), dict(weight='tensor', dual_clip=2.0), dict(pretrained=True, kl_type='k3', weight='tensor'),
          dict(use_value_clip=False, clip_ratio=0.05)][kind]
    op, t, p = cases.ppo_case(2000 + B + N, B, N, **kw)
    _check(op, t, p)


@pytest.mark.parametrize('B,N,nstep,kind', _draws(3, 14, range(1, 3000), [1, 2, 6, 18, 51], range(1, 8), [0, 1, 2]))
def test_fuzz_q_nstep(B, N, nstep, kind):
    kw = [dict(), dict(weight='tensor', value_gamma='tensor'), dict(rescale=True, done='bern')][kind]
    op, t, p = cases.qntd_case(3000 + B + N, B, N, nstep, **kw)
    _check(op, t, p)


@pytest.mark.parametrize('B,N,n_atom,nstep', _draws(4, 10, range(1, 900), [1, 3, 6], [2, 11, 51, 101], range(1, 6)))
def test_fuzz_c51(B, N, n_atom, nstep):
    op, t, p = cases.dntd_case(4000 + B + N, B, N, n_atom, nstep, weight='tensor' if B % 2 else 'none',
                               value_gamma='tensor' if B % 3 == 0 else 'none')
    _check(op, t, p)


@pytest.mark.parametrize('T,B,N', _draws(5, 14, range(1, 150), range(1, 500), [2, 3, 6, 7, 15, 18, 40]))
def test_fuzz_vtrace(T, B, N):
    op, t, p = cases.vtrace_case(5000 + T + B, T, B, N, weight='tensor' if T % 2 else 'none', rho_clip_ratio=0.9,
                                 c_clip_ratio=1.1, rho_pg_clip_ratio=1.2)
    _check(op, t, p)


@pytest.mark.parametrize('T,B', _draws(6, 10, range(1, 400), range(1, 300)))
def test_fuzz_td_lambda_and_returns(T, B):
    op, t, p = cases.td_lambda_case(6000 + T + B, T, B, weight='tensor' if B % 2 else 'none')
    _check(op, t, p)
    g = torch.Generator().manual_seed(T * 1000 + B)
    v, r = torch.randn(T + 1, B, generator=g), torch.randn(T, B, generator=g)
    gam, lam = torch.rand(T, B, generator=g), torch.rand(T, B, generator=g)
    done = (torch.rand(T, B, generator=g) < 0.1).float()
    want = rl_oracle.generalized_lambda_returns(v, r, gam, lam, done)
    got = b2.generalized_lambda_returns(v.to(DEV), r.to(DEV), gam.to(DEV), lam.to(DEV), done.to(DEV))
    assert torch.equal(got.cpu(), want)
    assert torch.equal(b2.upgo_returns(r.to(DEV), v.to(DEV)).cpu(), rl_oracle.upgo_returns(r, v))


@pytest.mark.parametrize('T,B,N', _draws(7, 8, range(1, 80), range(1, 120), [2, 6, 33, 100, 300]))
def test_fuzz_upgo(T, B, N):
    op, t, p = cases.upgo_case(7000 + T + B, T, B, N)
    _check(op, t, p)


@pytest.mark.parametrize('T,B,N', _draws(8, 8, range(1, 200), range(1, 300), [1, 2, 6, 18]))
def test_fuzz_retrace(T, B, N):
    op, t, p = cases.retrace_case(8000 + T + B, T, B, N)
    _check(op, t, p, exact=True)


@pytest.mark.parametrize('kind,B,N,n,n_p', _draws(9, 9, [0, 1, 2], range(1, 200), [1, 3, 6], [1, 8, 33, 200], [1, 8, 32, 150]))
def test_fuzz_quantile(kind, B, N, n, n_p):
    k = ['qrdqn', 'iqn', 'fqf'][kind]
    if k == 'qrdqn':
        n_p = n_p  # tau broadcast over the target axis works for any n'
    op, t, p = cases.quantile_case(9000 + B + n, k, B, N, n, n_p, 1 + B % 4, weight='tensor' if B % 2 else 'none',
                                   value_gamma='tensor' if B % 3 == 0 else 'none')
    _check(op, t, p)
