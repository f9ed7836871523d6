"""Seeded parity cases for the nine hot-path operators + adapters that run them through

  * an object exposing the **reference API** (``ding.rl_utils`` names/namedtuples): the real reference loaded by
    ``oracle/ref_loader.py`` or the product package ``di_engine_b200.rl_utils`` -- same adapter for both;
  * the flat-argument CPU oracle ``oracle/rl_oracle.py``.

A case = (op, tensors, params).  ``tensors`` holds CPU tensors (or None / python floats for the weight/value_gamma
variants); names listed in ``GRAD_INPUTS[op]`` get ``requires_grad``.  ``run_*`` return a flat dict of numpy arrays:
``out_*`` forward results, ``grad_*`` input gradients of ``sum_k c_k * loss_k`` with the fixed mixing coefficients
``LOSS_MIX`` (so every loss head's backward is exercised with a distinct upstream gradient).
"""
import copy
from collections import OrderedDict

import numpy as np
import torch

GRAD_INPUTS = {
    'gae': [],
    'ppo': ['logit_new', 'value_new'],
    'ppo_policy': ['logit_new'],
    'ppo_value': ['value_new'],
    'ppoc': ['mu_new', 'sigma_new', 'value_new'],
    'a2c': ['logit', 'value'],
    'vtc': ['mu_target', 'sigma_target', 'value'],
    'qntd': ['q'],
    'qntd_rescale': ['q'],
    'q1td': ['q'],
    'v1td': ['v'],
    'vntd': ['v'],
    'bdq': ['q'],
    'qseq': ['q'],
    'd1td': ['dist'],
    'dntd': ['dist'],
    'td_lambda': ['value'],
    'upgo': ['target_output'],
    'vtrace': ['target_output', 'value'],
    'qrdqn': ['q'],
    'iqn': ['q'],
    'fqf': ['q'],
    'retrace': [],
    'happo': ['logit_new', 'value_new'],
    'acer': ['target_logit', 'q_values'],
    'happoc': ['mu_new', 'sigma_new', 'value_new'],
    'ppg': ['logit_new', 'value_new'],
}
# upstream gradient for each returned loss head (distinct, non-trivial)
LOSS_MIX = {
    'ppo': [1.0, 0.5, -0.01, 0.3],
    'ppo_policy': [1.0, -0.02, 0.3],
    'ppo_value': [0.7],
    'ppoc': [1.0, 0.5, -0.01, 0.3],
    'a2c': [1.0, 0.5, -0.01],
    'vtc': [1.0, 0.5, -0.01],
    'qntd': [1.0],
    'qntd_rescale': [1.0],
    'q1td': [1.0],
    'v1td': [1.0],
    'vntd': [1.0],
    'bdq': [1.0],
    'qseq': [1.0],
    'd1td': [1.0],
    'dntd': [1.0],
    'td_lambda': [1.0],
    'upgo': [1.0],
    'vtrace': [1.0, 0.5, -0.01],
    'qrdqn': [1.0],
    'iqn': [1.0],
    'fqf': [1.0],
    'happo': [1.0, 0.5, -0.01],
    'happoc': [1.0, 0.5, -0.01],
    'ppg': [1.0, 0.7],
}


def _g(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _randn(g, *shape):
    return torch.randn(*shape, generator=g)


def _rand(g, *shape):
    return torch.rand(*shape, generator=g)


def _randint(g, hi, *shape):
    return torch.randint(0, hi, shape, generator=g)


def _bern(g, p, *shape):
    return (torch.rand(*shape, generator=g) < p).float()


# ----------------------------------------------------------------------------------------------------------------
# case builders
# ----------------------------------------------------------------------------------------------------------------
def gae_case(seed, T, B, A=None, done='float', traj='float', gamma=0.99, lambda_=0.97, p_done=0.05, one_d=False):
    g = _g(seed)
    shp = (T, ) if one_d else (T, B)
    vshp = shp if A is None else shp + (A, )
    t = OrderedDict()
    t['value'] = _randn(g, *vshp)
    t['next_value'] = _randn(g, *vshp)
    t['reward'] = _randn(g, *shp)
    d = _bern(g, p_done, *shp)
    tf = torch.clamp(d + _bern(g, p_done, *shp), max=1.0)
    tf[-1] = 1.0
    t['done'] = None if done is None else (d.bool() if done == 'bool' else d)
    t['traj_flag'] = None if traj is None else (tf.bool() if traj == 'bool' else tf)
    return 'gae', t, dict(gamma=gamma, lambda_=lambda_)


def ppo_case(seed, B, N, A=None, weight='none', pretrained=False, lead=None, **params):
    g = _g(seed)
    rows = (B, ) if A is None else (B, A)
    if lead is not None:
        rows = tuple(lead) + rows
    samp = rows if A is None else rows[:-1]
    t = OrderedDict()
    t['logit_new'] = _randn(g, *rows, N)
    t['logit_old'] = t['logit_new'] + 0.1 * _rand(g, *rows, N)
    t['action'] = _randint(g, N, *rows)
    t['value_new'] = _randn(g, *samp)
    t['value_old'] = t['value_new'] + 0.1 * _rand(g, *samp)
    t['adv'] = _randn(g, *samp)
    t['return_'] = _randn(g, *samp) * 2
    t['weight'] = None if weight == 'none' else _rand(g, *samp)
    t['logit_pretrained'] = (t['logit_new'] + 0.3 * _randn(g, *rows, N)) if pretrained else None
    return 'ppo', t, params


def qntd_case(seed, B, N, nstep, weight='none', value_gamma='none', gamma=0.95, cum_reward=False, list_gamma=False,
              marl_A=None, done='randn', rescale=False):
    g = _g(seed)
    t = OrderedDict()
    if marl_A is None:
        t['q'] = _randn(g, B, N)
        t['next_n_q'] = _randn(g, B, N)
        t['action'] = _randint(g, N, B)
        t['next_n_action'] = _randint(g, N, B)
    else:
        t['q'] = _randn(g, B, marl_A, N)
        t['next_n_q'] = _randn(g, B, marl_A, N)
        t['action'] = _randint(g, N, B, marl_A)
        t['next_n_action'] = _randint(g, N, B, marl_A)
    t['reward'] = _rand(g, B) if cum_reward else _rand(g, nstep, B)
    t['done'] = _randn(g, B) if done == 'randn' else _bern(g, 0.3, B)
    t['weight'] = None if weight == 'none' else _rand(g, B)
    params = dict(gamma=gamma, nstep=nstep)
    if not rescale:
        params['cum_reward'] = cum_reward
    if list_gamma:
        params['gamma'] = [torch.tensor(0.9 + 0.01 * i) for i in range(B)]
    if value_gamma == 'tensor':
        t['value_gamma'] = _rand(g, B)
    elif value_gamma == 'float':
        params['value_gamma'] = 0.857
    return ('qntd_rescale' if rescale else 'qntd'), t, params


def ppo_policy_case(seed, B, N, A=None, weight='none', pretrained=False, **params):
    op, t, _ = ppo_case(seed, B, N, A=A, weight=weight, pretrained=pretrained)
    keep = OrderedDict((k, t[k]) for k in ('logit_new', 'logit_old', 'action', 'adv', 'weight', 'logit_pretrained'))
    return 'ppo_policy', keep, params


def ppo_value_case(seed, B, weight='none', **params):
    op, t, _ = ppo_case(seed, B, 3, weight=weight)
    keep = OrderedDict((k, t[k]) for k in ('value_new', 'value_old', 'return_', 'weight'))
    return 'ppo_value', keep, params


def ppoc_case(seed, B, D, weight='none', pretrained=False, old_1d=False, **params):
    """ppo_error_continuous (tests/test_ppo.py:71-92): Independent(Normal(mu, sigma)) policies"""
    g = _g(seed)
    t = OrderedDict()
    t['mu_new'] = _rand(g, B, D)
    t['sigma_new'] = _rand(g, B, D) + 0.3
    if old_1d:
        assert D == 1
        t['mu_old'] = t['mu_new'][:, 0] + 0.1 * _rand(g, B)
        t['sigma_old'] = t['sigma_new'][:, 0] + 0.1 * _rand(g, B)
    else:
        t['mu_old'] = t['mu_new'] + 0.1 * _rand(g, B, D)
        t['sigma_old'] = t['sigma_new'] + 0.1 * _rand(g, B, D)
    t['action'] = _rand(g, B, D)
    t['value_new'] = _randn(g, B)
    t['value_old'] = t['value_new'] + 0.1 * _rand(g, B)
    t['adv'] = _randn(g, B)
    t['return_'] = _randn(g, B) * 2
    t['weight'] = None if weight == 'none' else _rand(g, B) + 1
    t['mu_pretrained'] = (t['mu_new'] + 0.3 * _randn(g, B, D)) if pretrained else None
    t['sigma_pretrained'] = (t['sigma_new'] + 0.2 * _rand(g, B, D)) if pretrained else None
    return 'ppoc', t, params


def a2c_case(seed, B, N, weight='none'):
    g = _g(seed)
    t = OrderedDict()
    t['logit'] = _randn(g, B, N)
    t['action'] = _randint(g, N, B)
    t['value'] = _randn(g, B)
    t['adv'] = _rand(g, B)
    t['return_'] = _randn(g, B) * 2
    t['weight'] = None if weight == 'none' else _rand(g, B) + 1
    return 'a2c', t, {}


def q1td_case(seed, B, N, weight='none', gamma=0.95):
    g = _g(seed)
    t = OrderedDict()
    t['q'] = _randn(g, B, N)
    t['next_q'] = _randn(g, B, N)
    t['act'] = _randint(g, N, B)
    t['next_act'] = _randint(g, N, B)
    t['reward'] = _rand(g, B)
    t['done'] = _bern(g, 0.3, B)
    t['weight'] = None if weight == 'none' else _rand(g, B)
    return 'q1td', t, dict(gamma=gamma)


def v1td_case(seed, B, K=None, weight='none', done='bern', gamma=0.95):
    g = _g(seed)
    shape = (B, ) if K is None else (B, K)
    t = OrderedDict()
    t['v'] = _randn(g, *shape)
    t['next_v'] = _randn(g, *shape)
    t['reward'] = _rand(g, B)
    t['done'] = None if done == 'none' else _bern(g, 0.3, B)
    t['weight'] = None if weight == 'none' else _rand(g, *shape)
    return 'v1td', t, dict(gamma=gamma)


def vntd_case(seed, B, nstep, weight='none', value_gamma='none', gamma=0.95):
    g = _g(seed)
    t = OrderedDict()
    t['v'] = _randn(g, B)
    t['next_n_v'] = _randn(g, B)
    t['reward'] = _rand(g, nstep, B)
    t['done'] = _bern(g, 0.3, B)
    t['weight'] = None if weight == 'none' else _rand(g, B)
    t['value_gamma'] = None if value_gamma == 'none' else _rand(g, B)
    return 'vntd', t, dict(gamma=gamma, nstep=nstep)


def bdq_case(seed, B, D, N, nstep, weight='none', value_gamma='none', gamma=0.95, cum_reward=False):
    g = _g(seed)
    t = OrderedDict()
    t['q'] = _randn(g, B, D, N)
    t['next_n_q'] = _randn(g, B, D, N)
    t['action'] = _randint(g, N, B, D)
    t['next_n_action'] = _randint(g, N, B, D)
    t['reward'] = _rand(g, B) if cum_reward else _rand(g, nstep, B)
    t['done'] = _bern(g, 0.3, B)
    t['weight'] = None if weight == 'none' else _rand(g, B)
    params = dict(gamma=gamma, nstep=nstep, cum_reward=cum_reward)
    if value_gamma == 'tensor':
        t['value_gamma'] = _rand(g, B)
    return 'bdq', t, params


def qseq_case(seed, T, B, N, nstep, weight='tensor', value_gamma='tensor', gamma=0.997, rescale=False, list_gamma=False):
    """The recurrent learners' per-step loop (ding/policy/r2d2.py:347-369): reward already in the (T, nstep, B) layout."""
    g = _g(seed)
    t = OrderedDict()
    t['q'] = _randn(g, T, B, N)
    t['next_n_q'] = _randn(g, T, B, N)
    t['action'] = _randint(g, N, T, B)
    t['next_n_action'] = _randint(g, N, T, B)
    t['reward'] = _rand(g, T, nstep, B)
    t['done'] = _bern(g, 0.1, T, B)
    t['weight'] = None if weight == 'none' else _rand(g, T, B)
    t['value_gamma'] = None if value_gamma == 'none' else _rand(g, T, B)
    params = dict(gamma=gamma, nstep=nstep, rescale=rescale)
    if list_gamma:
        params['gamma'] = [torch.tensor(0.9 + 0.01 * i) for i in range(B)]
    return 'qseq', t, params


def d1td_case(seed, B, N, n_atom, gamma=0.95, v_min=-10., v_max=10., marl_A=None):
    g = _g(seed)
    lead = (B, ) if marl_A is None else (B, marl_A)
    t = OrderedDict()
    t['dist'] = torch.softmax(_randn(g, *lead, N, n_atom), -1)
    t['next_dist'] = torch.softmax(_randn(g, *lead, N, n_atom), -1)
    t['act'] = _randint(g, N, *lead)
    t['next_act'] = _randint(g, N, *lead)
    t['reward'] = _randn(g, B)
    t['done'] = _bern(g, 0.3, B)
    t['weight'] = None
    return 'd1td', t, dict(gamma=gamma, v_min=v_min, v_max=v_max, n_atom=n_atom)


def dntd_case(seed, B, N, n_atom, nstep, weight='none', value_gamma='none', gamma=0.95, v_min=-10., v_max=10.,
              marl_A=None, integer_bins=False, done='bern'):
    g = _g(seed)
    lead = (B, ) if marl_A is None else (B, marl_A)
    t = OrderedDict()
    t['dist'] = torch.softmax(_randn(g, *lead, N, n_atom), -1)
    t['next_n_dist'] = torch.softmax(_randn(g, *lead, N, n_atom), -1)
    t['act'] = _randint(g, N, *lead)
    t['next_n_act'] = _randint(g, N, *lead)
    t['reward'] = _randn(g, nstep, B)
    t['done'] = _bern(g, 0.3, B) if done == 'bern' else _randn(g, B)
    if integer_bins:
        # reward 0 and done 1 -> target_z == support*0 + 0 -> b is an exact integer for every atom of those rows;
        # large |reward| rows clamp to v_min / v_max (b == 0 or n_atom-1 exactly): exercises the l==u fix-ups.
        t['reward'] = torch.zeros(nstep, B)
        t['reward'][0, ::3] = 100.
        t['reward'][0, 1::3] = -100.
        t['done'] = torch.ones(B)
        t['done'][::2] = 0.
    params = dict(gamma=gamma, v_min=v_min, v_max=v_max, n_atom=n_atom, nstep=nstep)
    if weight == 'tensor':
        t['weight'] = _rand(g, B) if marl_A is None else _rand(g, B * marl_A)
    elif weight == 'one':
        t['weight'] = _rand(g, 1)
    elif weight == 'float':
        t['weight'] = None
        params['weight_float'] = 0.7
    else:
        t['weight'] = None
    if value_gamma == 'tensor':
        t['value_gamma'] = _rand(g, B)
    elif value_gamma == 'scalar_tensor':
        t['value_gamma'] = torch.tensor(0.9)
    elif value_gamma == 'float':
        params['value_gamma'] = 0.857
    return 'dntd', t, params


def td_lambda_case(seed, T, B, weight='none', gamma=0.9, lambda_=0.8):
    g = _g(seed)
    t = OrderedDict()
    t['value'] = _randn(g, T + 1, B)
    t['reward'] = _rand(g, T, B)
    t['weight'] = None if weight == 'none' else _rand(g, T, B)
    return 'td_lambda', t, dict(gamma=gamma, lambda_=lambda_)


def upgo_case(seed, T, B, N, N2=None, mask=False):
    g = _g(seed)
    t = OrderedDict()
    if N2 is None:
        t['target_output'] = _randn(g, T, B, N)
        t['action'] = _randint(g, N, T, B)
    else:
        t['target_output'] = _randn(g, T, B, N2, N)
        t['action'] = _randint(g, N, T, B, N2)
    t['rhos'] = _rand(g, T, B) + 0.5
    t['rewards'] = _randn(g, T, B)
    t['bootstrap_values'] = _randn(g, T + 1, B)
    t['mask'] = (_rand(g, T, B, N2) > 0.3).float() if mask else None
    return 'upgo', t, {}


def vtrace_case(seed, T, B, N, weight='none', **params):
    g = _g(seed)
    t = OrderedDict()
    t['target_output'] = _randn(g, T, B, N)
    t['behaviour_output'] = t['target_output'] + 0.5 * _randn(g, T, B, N)
    t['action'] = _randint(g, N, T, B)
    t['value'] = _randn(g, T + 1, B)
    t['reward'] = _rand(g, T, B)
    t['weight'] = None if weight == 'none' else _rand(g, T, B)
    return 'vtrace', t, params


def _qntd_cum_nb(seed):
    """cum_reward=True with an (nstep, B) reward: the reference broadcasts to (nstep, B) errors (tests/test_td.py:29-35)."""
    g = _g(seed)
    B, N, nstep = 6, 4, 3
    t = OrderedDict()
    t['q'] = _randn(g, B, N)
    t['next_n_q'] = _randn(g, B, N)
    t['action'] = _randint(g, N, B)
    t['next_n_action'] = _randint(g, N, B)
    t['reward'] = _rand(g, nstep, B)
    t['done'] = _randn(g, B)
    t['weight'] = None
    t['value_gamma'] = torch.tensor(0.9)
    return t, dict(gamma=0.95, nstep=nstep, cum_reward=True)


def _qntd_marl(seed):
    """The reference's multi-agent branch: action (B, A, 1) against q (B, A, N) (td.py:700-705)."""
    g = _g(seed)
    B, A, N, nstep = 5, 3, 4, 2
    t = OrderedDict()
    t['q'] = _randn(g, B, A, N)
    t['next_n_q'] = _randn(g, B, A, N)
    t['action'] = _randint(g, N, B, A, 1)
    t['next_n_action'] = _randint(g, N, B, A)
    t['reward'] = _rand(g, nstep, B)
    t['done'] = _bern(g, 0.3, B)
    t['weight'] = _rand(g, B)
    t['value_gamma'] = _rand(g, B)
    return t, dict(gamma=0.9, nstep=nstep, cum_reward=False)


def vtc_case(seed, T, B, D, weight='none', **params):
    """vtrace_error_continuous_action (tests/test_vtrace.py:25-47)"""
    g = _g(seed)
    t = OrderedDict()
    t['mu_target'] = _randn(g, T, B, D)
    t['sigma_target'] = torch.exp(0.3 * _randn(g, T, B, D))
    t['mu_behaviour'] = t['mu_target'] + 0.2 * _randn(g, T, B, D)
    t['sigma_behaviour'] = torch.exp(0.3 * _randn(g, T, B, D))
    t['action'] = _randn(g, T, B, D)
    t['value'] = _randn(g, T + 1, B)
    t['reward'] = _rand(g, T, B)
    t['weight'] = None if weight == 'none' else _rand(g, T, B)
    return 'vtc', t, params


def happo_case(seed, B, N, weight='none', A=None, **params):
    """ppo_case + the per-sample factor (B, 1) of happo_data (happo.py:12-14); A: (B, A, N) logits against (B,) adv, the
    ratio.mean(dim=1) branch (happo.py:114-121)"""
    op, t, p = ppo_case(seed, B, N, A=A, weight=weight, **params)
    t = OrderedDict((k, v) for k, v in t.items() if k != 'logit_pretrained')
    t['factor'] = _rand(_g(seed + 7919), B, 1) * 1.5 + 0.25
    return 'happo', t, p


def acer_case(seed, T, B, N, c_clip_ratio=10.0, trust_region_value=1.0):
    """the operands of ACERPolicy._forward_learn (policy/acer.py:215-260): log-softmax policy outputs, Retrace targets, ratios
    on both sides of the truncation, an actor gradient for the trust-region projection"""
    g = _g(seed)
    t = OrderedDict()
    t['q_values'] = _randn(g, T, B, N)
    t['q_retraces'] = _randn(g, T, B, 1)
    t['v_pred'] = _randn(g, T, B, 1)
    t['target_logit'] = torch.log_softmax(_randn(g, T, B, N), dim=-1)
    t['actions'] = _randint(g, N, T, B)
    t['ratio'] = _rand(g, T, B, N) * 3.0 * c_clip_ratio / 2.0  # both sides of 1 - c / ratio > 0
    t['avg_logit'] = torch.log_softmax(_randn(g, T, B, N), dim=-1)
    t['actor_gradient'] = _randn(g, T, B, N)
    return 'acer', t, dict(c_clip_ratio=c_clip_ratio, trust_region_value=trust_region_value)


def happoc_case(seed, B, D, weight='none', **params):
    """happo_error_continuous (tests/test_happo.py:49-75): ppoc_case + the per-sample factor"""
    op, t, p = ppoc_case(seed, B, D, weight=weight, **params)
    t = OrderedDict((k, v) for k, v in t.items() if not k.endswith('_pretrained'))
    t['factor'] = _rand(_g(seed + 7919), B, 1) * 1.5 + 0.25
    return 'happoc', t, p


def ppg_case(seed, B, N, weight='none', **params):
    """ppg_joint_error (tests/test_ppg.py): ppo_case without adv / pretrained"""
    op, t, p = ppo_case(seed, B, N, weight=weight, **params)
    t = OrderedDict((k, t[k]) for k in ('logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'return_', 'weight'))
    return 'ppg', t, p


def retrace_case(seed, T, B, N, gamma=0.99):
    """tests/test_retrace.py:8-18"""
    g = _g(seed)
    t = OrderedDict()
    t['q_values'] = _randn(g, T + 1, B, N)
    t['v_pred'] = _randn(g, T + 1, B, 1)
    t['rewards'] = _randn(g, T, B)
    t['actions'] = _randint(g, N, T, B)
    t['weights'] = _rand(g, T, B)
    t['ratio'] = _rand(g, T, B, N) * 0.8 + 0.6  # both sides of the clamp at 1
    return 'retrace', t, dict(gamma=gamma)


HAPPO_FIELDS = ('logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight', 'factor')
_QF = ('q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done')
QUANTILE_FIELDS = {'qrdqn': _QF + ('tau', 'weight'), 'iqn': _QF + ('replay_quantiles', 'weight'), 'fqf': _QF + ('quantiles_hats', 'weight')}


def quantile_case(seed, kind, B, N, n, n_p, nstep, weight='none', value_gamma='none', gamma=0.95, kappa=None, tau='tensor'):
    """kind 'qrdqn' (q (B, N, n)), 'iqn' (q (n, B, N), replay quantiles (n, B, 1)) or 'fqf' (q (B, n, N), quantile hats (B, n))"""
    g = _g(seed)
    t = OrderedDict()
    shape = {'qrdqn': lambda k: (B, N, k), 'iqn': lambda k: (k, B, N), 'fqf': lambda k: (B, k, N)}[kind]
    t['q'] = _randn(g, *shape(n))
    t['next_n_q'] = _randn(g, *shape(n_p))
    t['action'] = _randint(g, N, B)
    t['next_n_action'] = _randint(g, N, B)
    t['reward'] = _rand(g, nstep, B)
    t['done'] = _bern(g, 0.3, B)
    if kind == 'qrdqn':
        mid = (torch.arange(n, dtype=torch.float32) + 0.5) / n  # the QRDQN head's quantile midpoints
        t['tau'] = {'tensor': mid.view(1, n, 1).repeat(B, 1, 1), 'row': mid.view(1, n, 1), 'scalar': torch.tensor(0.3)}[tau]
    elif kind == 'iqn':
        t['replay_quantiles'] = _rand(g, n, B, 1)
    else:
        t['quantiles_hats'] = _rand(g, B, n).sort(dim=1).values
    t['weight'] = None if weight == 'none' else _rand(g, B) + 0.5
    p = dict(gamma=gamma, nstep=nstep)
    if value_gamma != 'none':
        t['value_gamma'] = torch.tensor(0.9) if value_gamma == 'scalar' else _rand(g, B) * 0.5 + 0.5
    if kappa is not None:
        p['kappa'] = kappa
    return kind, t, p


def build_cases():
    """Small cases: what the golden fixtures hold and what every implementation is compared on."""
    c = OrderedDict()
    # ---- gae (configs A + reference test shapes tests/test_gae.py:7-36) -------------------------------------
    c['gae_cfgA'] = gae_case(1, 128, 8, gamma=0.9, lambda_=0.95)
    c['gae_none'] = gae_case(2, 32, 6, done=None, traj=None)
    c['gae_done_only'] = gae_case(3, 32, 6, traj=None, p_done=0.2)
    c['gae_bool'] = gae_case(4, 40, 5, done='bool', traj='bool', p_done=0.2)
    c['gae_1d'] = gae_case(5, 257, 1, one_d=True, p_done=0.03)
    c['gae_marl'] = gae_case(6, 24, 4, A=3, p_done=0.2)
    c['gae_wide'] = gae_case(7, 16, 200, p_done=0.1)
    c['gae_T1'] = gae_case(8, 1, 7)
    # ---- ppo (tests/test_ppo.py:24-92) -----------------------------------------------------------------------
    c['ppo_cfgA'] = ppo_case(10, 64, 2, clip_ratio=0.2)
    i = 11
    for uvc in (True, False):
        for dc in (None, 5.0):
            for w in ('none', 'tensor'):
                c['ppo_vc%d_dc%d_w%d' % (uvc, dc is not None, w == 'tensor')] = ppo_case(
                    i, 48, 6, weight=w, use_value_clip=uvc, dual_clip=dc, clip_ratio=0.2
                )
                i += 1
    for k in ('k1', 'k2', 'k3'):
        c['ppo_kl_' + k] = ppo_case(i, 33, 5, pretrained=True, kl_type=k, weight='tensor')
        i += 1
    c['ppo_marl'] = ppo_case(i, 12, 7, A=4, weight='tensor')
    c['ppo_marl_dc'] = ppo_case(i + 1, 12, 7, A=4, dual_clip=3.0)
    c['ppo_wideN'] = ppo_case(i + 2, 9, 130, weight='tensor', clip_ratio=0.1)
    c['ppo_seq'] = ppo_case(i + 3, 5, 6, lead=(3, ), weight='tensor')
    c['ppo_one'] = ppo_case(i + 4, 1, 3)
    c['ppo_policy_basic'] = ppo_policy_case(i + 5, 64, 6, clip_ratio=0.2)
    c['ppo_policy_w_dc_kl'] = ppo_policy_case(i + 6, 33, 5, weight='tensor', pretrained=True, dual_clip=3.0, kl_type='k3')
    c['ppo_policy_marl_noent'] = ppo_policy_case(i + 7, 12, 7, A=4, entropy_bonus=False)
    c['ppo_value_clip'] = ppo_value_case(i + 8, 100, weight='tensor', clip_ratio=0.2)
    c['ppo_value_noclip'] = ppo_value_case(i + 9, 37, use_value_clip=False)
    # ---- q_nstep (tests/test_td.py:13-126) -------------------------------------------------------------------
    c['qntd_cfgB'] = qntd_case(30, 64, 6, 3, value_gamma='tensor', gamma=0.99, done='bern')
    c['qntd_n1'] = qntd_case(31, 5, 4, 1)
    c['qntd_n5_w'] = qntd_case(32, 17, 3, 5, weight='tensor')
    c['qntd_cum'] = qntd_case(33, 17, 3, 5, cum_reward=True, value_gamma='tensor')
    c['qntd_cum_novg'] = qntd_case(34, 17, 3, 5, cum_reward=True)
    c['qntd_vg_float'] = qntd_case(35, 9, 3, 2, value_gamma='float')
    c['qntd_ngu'] = qntd_case(36, 6, 4, 3, list_gamma=True)
    c['qntdr_n3'] = qntd_case(40, 33, 6, 3, rescale=True)
    c['qntdr_n5_w_vg'] = qntd_case(41, 9, 4, 5, rescale=True, weight='tensor', value_gamma='tensor')
    c['qntdr_ngu'] = qntd_case(42, 6, 4, 3, rescale=True, list_gamma=True)
    # ---- 1-step / state-value siblings on the q-n-step kernels (tests/test_td.py:207-268) -----------------------
    c['q1td_basic'] = q1td_case(43, 12, 5)
    c['q1td_w'] = q1td_case(44, 64, 6, weight='tensor', gamma=0.99)
    c['v1td_basic'] = v1td_case(45, 16, gamma=0.99)
    c['v1td_w_nodone'] = v1td_case(46, 9, weight='tensor', done='none')
    c['v1td_2d'] = v1td_case(47, 8, K=3, weight='tensor')
    c['vntd_n3'] = vntd_case(48, 10, 3, gamma=0.99)
    c['vntd_n5_w_vg'] = vntd_case(49, 7, 5, weight='tensor', value_gamma='tensor')
    # ---- shapes beyond (B, N)/(B,) that the reference's broadcasting accepts (tests/test_td.py:29-35, td.py:700-705) ----
    c['qntd_cum_nstep_reward'] = ('qntd', ) + _qntd_cum_nb(90)
    c['qntd_marl_branch'] = ('qntd', ) + _qntd_marl(91)
    # ---- bdq_nstep (tests/test_td.py:40-68), the recurrent learners' loop (policy/r2d2.py:347-369), dist_1step -------------
    c['bdq_n3'] = bdq_case(92, 8, 6, 3, 3)
    c['bdq_n5_w_vg'] = bdq_case(93, 9, 4, 5, 5, weight='tensor', value_gamma='tensor')
    c['bdq_cum'] = bdq_case(94, 7, 3, 4, 2, cum_reward=True, weight='tensor')
    c['qseq_r2d2'] = qseq_case(95, 10, 12, 6, 3)
    c['qseq_rescale_now'] = qseq_case(96, 7, 5, 4, 2, weight='none', value_gamma='none', rescale=True)
    c['qseq_ngu'] = qseq_case(97, 5, 6, 3, 2, list_gamma=True, value_gamma='none')
    c['d1td_basic'] = d1td_case(98, 9, 4, 51)
    c['d1td_marl'] = d1td_case(99, 4, 3, 21, marl_A=2, v_min=-2., v_max=3.)
    # ---- sibling heads: ppo_error_continuous (tests/test_ppo.py:71-92), a2c_error (tests/test_a2c.py)
    c['ppoc_basic'] = ppoc_case(110, 64, 6, clip_ratio=0.2)
    c['ppoc_w_dc_novc'] = ppoc_case(111, 33, 3, weight='tensor', dual_clip=5.0, use_value_clip=False)
    c['ppoc_kl_k3'] = ppoc_case(112, 20, 4, weight='tensor', pretrained=True, kl_type='k3')
    c['ppoc_old_1d'] = ppoc_case(113, 17, 1, old_1d=True)
    c['vtc_small'] = vtc_case(116, 4, 8, 16, rho_clip_ratio=1.1)
    c['vtc_w_clips'] = vtc_case(117, 13, 5, 3, weight='tensor', rho_clip_ratio=0.8, c_clip_ratio=1.3, rho_pg_clip_ratio=2.0)
    # ---- quantile-regression TD heads (tests/test_td.py:249-268, :461-505) -----------------------------------------------------
    c['qrdqn_basic'] = quantile_case(130, 'qrdqn', 8, 4, 7, 7, 3)
    c['qrdqn_w_vg'] = quantile_case(131, 'qrdqn', 6, 3, 32, 32, 5, weight='tensor', value_gamma='tensor', tau='row')
    c['qrdqn_scalar_tau'] = quantile_case(132, 'qrdqn', 4, 3, 3, 3, 1, value_gamma='scalar', tau='scalar')
    c['iqn_basic'] = quantile_case(133, 'iqn', 8, 4, 8, 6, 3)
    c['iqn_w_kappa'] = quantile_case(134, 'iqn', 5, 6, 32, 32, 2, weight='tensor', value_gamma='tensor', kappa=0.6)
    c['fqf_basic'] = quantile_case(135, 'fqf', 8, 4, 8, 8, 3)
    c['fqf_w_kappa'] = quantile_case(136, 'fqf', 5, 6, 32, 16, 4, weight='tensor', value_gamma='scalar', kappa=1.7)
    # ---- ACER heads (no reference unit test; operands as policy/acer.py builds them) -------------------------------------------
    c['acer_small'] = acer_case(160, 5, 4, 6, c_clip_ratio=2.0, trust_region_value=0.2)
    c['acer_default'] = acer_case(161, 16, 9, 3)
    # ---- HAPPO (tests/test_happo.py) -------------------------------------------------------------------------------------------
    c['happo_basic'] = happo_case(150, 64, 6, clip_ratio=0.2)
    c['happo_w_dc'] = happo_case(151, 33, 5, weight='tensor', dual_clip=3.0, clip_ratio=0.3)
    c['happo_noclip'] = happo_case(152, 12, 40, use_value_clip=False)
    c['happo_marl'] = happo_case(157, 12, 7, A=4, weight='tensor', dual_clip=3.0)
    c['happoc_basic'] = happoc_case(153, 16, 6)
    c['ppg_basic'] = ppg_case(155, 32, 6)
    c['ppg_w_noclip'] = ppg_case(156, 9, 17, weight='tensor', use_value_clip=False, clip_ratio=0.1)
    c['happoc_w_dc'] = happoc_case(154, 33, 3, weight='tensor', dual_clip=3.0, clip_ratio=0.3, use_value_clip=False)
    # ---- ACER Retrace targets (tests/test_retrace.py) ----------------------------------------------------------------------
    c['retrace_ref_test'] = retrace_case(140, 64, 32, 6)
    c['retrace_ragged'] = retrace_case(141, 13, 5, 3, gamma=0.9)
    c['a2c_basic'] = a2c_case(114, 64, 6)
    c['a2c_w_wide'] = a2c_case(115, 9, 130, weight='tensor')
    # ---- dist_nstep (tests/test_td.py:130-204) ---------------------------------------------------------------
    c['dntd_cfgC'] = dntd_case(50, 32, 6, 51, 3, gamma=0.99, value_gamma='tensor')
    c['dntd_n5'] = dntd_case(51, 4, 3, 51, 5)
    c['dntd_w_tensor'] = dntd_case(52, 7, 3, 51, 5, weight='tensor')
    c['dntd_w_one'] = dntd_case(53, 7, 3, 51, 2, weight='one', done='randn')
    c['dntd_w_float'] = dntd_case(54, 7, 3, 21, 2, weight='float', value_gamma='float')
    c['dntd_vg_scalar'] = dntd_case(55, 7, 3, 51, 5, value_gamma='scalar_tensor')
    c['dntd_intbins'] = dntd_case(56, 12, 4, 51, 2, integer_bins=True, gamma=1.0)
    c['dntd_marl'] = dntd_case(57, 4, 3, 51, 5, marl_A=2)
    c['dntd_atoms200'] = dntd_case(58, 5, 2, 200, 3, v_min=-1., v_max=5.)
    # ---- td_lambda / upgo / vtrace ---------------------------------------------------------------------------
    c['tdl_basic'] = td_lambda_case(60, 8, 4)
    c['tdl_w'] = td_lambda_case(61, 33, 17, weight='tensor', gamma=0.99, lambda_=0.95)
    c['tdl_T1'] = td_lambda_case(62, 1, 5)
    c['upgo_3d'] = upgo_case(70, 4, 8, 5)
    c['upgo_3d_big'] = upgo_case(71, 33, 19, 6)
    c['upgo_4d'] = upgo_case(72, 4, 8, 5, N2=7)
    c['upgo_4d_mask'] = upgo_case(73, 4, 8, 5, N2=7, mask=True)
    c['vtrace_small'] = vtrace_case(80, 4, 8, 16, rho_clip_ratio=1.1)
    c['vtrace_cfgE'] = vtrace_case(81, 16, 24, 6, weight='tensor', gamma=0.99, lambda_=0.95)
    c['vtrace_clips'] = vtrace_case(82, 9, 5, 3, rho_clip_ratio=0.8, c_clip_ratio=1.3, rho_pg_clip_ratio=2.0)
    return c


# ----------------------------------------------------------------------------------------------------------------
# adapters
# ----------------------------------------------------------------------------------------------------------------
def prepare(op, tensors, device='cpu'):
    """Deep-copy the case tensors to ``device`` and switch on requires_grad for the differentiable inputs."""
    out = OrderedDict()
    for k, v in tensors.items():
        if isinstance(v, torch.Tensor):
            v = v.clone().to(device)
            if k in GRAD_INPUTS[op]:
                v.requires_grad_(True)
        out[k] = v
    return out


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy()
    return np.asarray(x)


def _backward(op, losses, t, res):
    mix = LOSS_MIX.get(op)
    if not mix:
        return
    total = None
    for c, l in zip(mix, losses):
        if isinstance(l, torch.Tensor) and l.requires_grad:
            total = c * l if total is None else total + c * l
    if total is None:
        return
    total.backward()
    for k in GRAD_INPUTS[op]:
        if t[k].grad is not None:
            res['grad_' + k] = _np(t[k].grad)


def run_api(api, op, tensors, params, device='cpu'):
    """Run a case through an object with the reference's names (the real reference or the product package)."""
    t = prepare(op, tensors, device)
    p = copy.copy(params)
    res = OrderedDict()
    if op == 'gae':
        data = api.gae_data(t['value'], t['next_value'], t['reward'], t['done'], t['traj_flag'])
        adv = api.gae(data, **p)
        res['out_adv'] = _np(adv)
        res['out_next_value_after'] = _np(t['next_value'])  # in-place mask, gae.py:61
        return res
    if op == 'ppo':
        data = api.ppo_data(*[t[k] for k in ('logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'adv',
                                              'return_', 'weight', 'logit_pretrained')])
        loss, info = api.ppo_error(data, **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss', 'kl_div'):
            res['out_' + k] = _np(getattr(loss, k))
        res['out_approx_kl'] = np.float32(info.approx_kl)
        res['out_clipfrac'] = np.float32(info.clipfrac)
        _backward(op, list(loss), t, res)
        return res
    if op == 'ppoc':
        pre = None if t['mu_pretrained'] is None else {'mu': t['mu_pretrained'], 'sigma': t['sigma_pretrained']}
        data = api.ppo_data({'mu': t['mu_new'], 'sigma': t['sigma_new']}, {'mu': t['mu_old'], 'sigma': t['sigma_old']},
                            t['action'], t['value_new'], t['value_old'], t['adv'], t['return_'], t['weight'], pre)
        loss, info = api.ppo_error_continuous(data, **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss', 'kl_div'):
            res['out_' + k] = _np(getattr(loss, k))
        res['out_approx_kl'] = np.float32(info.approx_kl)
        res['out_clipfrac'] = np.float32(info.clipfrac)
        _backward(op, list(loss), t, res)
        return res
    if op == 'vtc':
        data = api.vtrace_data({'mu': t['mu_target'], 'sigma': t['sigma_target']},
                               {'mu': t['mu_behaviour'], 'sigma': t['sigma_behaviour']}, t['action'], t['value'], t['reward'],
                               t['weight'])
        loss = api.vtrace_error_continuous_action(data, **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss'):
            res['out_' + k] = _np(getattr(loss, k))
        _backward(op, list(loss), t, res)
        return res
    if op == 'a2c':
        loss = api.a2c_error(api.a2c_data(t['logit'], t['action'], t['value'], t['adv'], t['return_'], t['weight']))
        for k in ('policy_loss', 'value_loss', 'entropy_loss'):
            res['out_' + k] = _np(getattr(loss, k))
        _backward(op, list(loss), t, res)
        return res
    if op == 'retrace':
        res['out_q_retraces'] = _np(api.compute_q_retraces(*t.values(), **p))
        return res
    if op == 'acer':
        actor, bias = api.acer_policy_error(t['q_values'].detach(), t['q_retraces'], t['v_pred'], t['target_logit'], t['actions'],
                                            t['ratio'], p['c_clip_ratio'])
        critic = api.acer_value_error(t['q_values'], t['q_retraces'], t['actions'])
        res['out_actor_loss'], res['out_bias_correction_loss'], res['out_critic_loss'] = _np(actor), _np(bias), _np(critic)
        w = torch.linspace(0.5, 1.5, actor.numel(), device=actor.device).reshape(actor.shape)
        ((actor * w).sum() + 0.3 * (bias * w.flip(0)).sum()).backward()
        (critic * w).sum().backward()
        res['grad_target_logit'], res['grad_q_values'] = _np(t['target_logit'].grad), _np(t['q_values'].grad)
        res['out_trust_region'] = _np(api.acer_trust_region_update([t['actor_gradient']], t['target_logit'].detach(),
                                                                   t['avg_logit'], p['trust_region_value'])[0])
        return res
    if op == 'ppg':
        loss = api.ppg_joint_error(api.ppg_data(*t.values()), **p)
        res['out_auxiliary_loss'], res['out_behavioral_cloning_loss'] = _np(loss[0]), _np(loss[1])
        _backward(op, list(loss), t, res)
        return res
    if op == 'happoc':
        data = api.happo_data({'mu': t['mu_new'], 'sigma': t['sigma_new']}, {'mu': t['mu_old'], 'sigma': t['sigma_old']},
                              t['action'], t['value_new'], t['value_old'], t['adv'], t['return_'], t['weight'], t['factor'])
        loss, info = api.happo_error_continuous(data, **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss'):
            res['out_' + k] = _np(getattr(loss, k))
        res['out_approx_kl'] = np.float32(info.approx_kl)
        res['out_clipfrac'] = np.float32(info.clipfrac)
        _backward(op, list(loss), t, res)
        return res
    if op == 'happo':
        loss, info = api.happo_error(api.happo_data(*[t[k] for k in HAPPO_FIELDS]), **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss'):
            res['out_' + k] = _np(getattr(loss, k))
        res['out_approx_kl'] = np.float32(info.approx_kl)
        res['out_clipfrac'] = np.float32(info.clipfrac)
        _backward(op, list(loss), t, res)
        return res
    if op in ('qrdqn', 'iqn', 'fqf'):
        data = getattr(api, op + '_nstep_td_data')(*[t[k] for k in QUANTILE_FIELDS[op]])
        loss, per = getattr(api, op + '_nstep_td_error')(data, value_gamma=t.get('value_gamma'), **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'ppo_policy':
        data = api.ppo_policy_data(*[t[k] for k in ('logit_new', 'logit_old', 'action', 'adv', 'weight',
                                                     'logit_pretrained')])
        loss, info = api.ppo_policy_error(data, **p)
        for k in ('policy_loss', 'entropy_loss', 'kl_div'):
            res['out_' + k] = _np(getattr(loss, k))
        res['out_approx_kl'] = np.float32(info.approx_kl)
        res['out_clipfrac'] = np.float32(info.clipfrac)
        _backward(op, list(loss), t, res)
        return res
    if op == 'ppo_value':
        loss = api.ppo_value_error(api.ppo_value_data(t['value_new'], t['value_old'], t['return_'], t['weight']), **p)
        res['out_value_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op in ('qntd', 'qntd_rescale'):
        data = api.q_nstep_td_data(*[t[k] for k in ('q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done',
                                                    'weight')])
        gamma = p.pop('gamma')
        if isinstance(gamma, list):
            gamma = [x.to(device) for x in gamma]
        if 'value_gamma' in t:
            p['value_gamma'] = t['value_gamma']
        fn = api.q_nstep_td_error if op == 'qntd' else api.q_nstep_td_error_with_rescale
        loss, per = fn(data, gamma, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'bdq':
        data = api.q_nstep_td_data(*[t[k] for k in ('q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done',
                                                    'weight')])
        if 'value_gamma' in t:
            p['value_gamma'] = t['value_gamma']
        loss, per = api.bdq_nstep_td_error(data, p.pop('gamma'), **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'qseq':
        gamma = p['gamma']
        if isinstance(gamma, list):
            gamma = [x.to(device) for x in gamma]
        if hasattr(api, 'q_nstep_td_error_sequence'):
            data = api.q_nstep_td_seq_data(*[t[k] for k in ('q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done',
                                                            'weight')])
            loss, prio, per = api.q_nstep_td_error_sequence(data, gamma, p['nstep'], value_gamma=t['value_gamma'],
                                                            rescale=p['rescale'])
        else:  # the reference: the loop of ding/policy/r2d2.py:347-369 around its operator
            fn = api.q_nstep_td_error_with_rescale if p['rescale'] else api.q_nstep_td_error
            losses, errs, raw = [], [], []
            for i in range(t['q'].shape[0]):
                td_data = api.q_nstep_td_data(t['q'][i], t['next_n_q'][i], t['action'][i], t['next_n_action'][i],
                                              t['reward'][i], t['done'][i],
                                              None if t['weight'] is None else t['weight'][i])
                l, e = fn(td_data, gamma, p['nstep'],
                          value_gamma=None if t['value_gamma'] is None else t['value_gamma'][i])
                losses.append(l)
                errs.append(e.abs())
                raw.append(e)
            loss = sum(losses) / (len(losses) + 1e-8)
            prio = 0.9 * torch.max(torch.stack(errs), dim=0)[0] + (1 - 0.9) * (torch.sum(torch.stack(errs), dim=0) /
                                                                               (len(errs) + 1e-8))
            per = torch.stack(raw)
        res['out_loss'] = _np(loss)
        res['out_priority'] = _np(prio)
        res['out_td_error'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'd1td':
        data = api.dist_1step_td_data(t['dist'], t['next_dist'], t['act'], t['next_act'], t['reward'], t['done'],
                                      t['weight'])
        loss = api.dist_1step_td_error(data, **p)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'q1td':
        data = api.q_1step_td_data(*[t[k] for k in ('q', 'next_q', 'act', 'next_act', 'reward', 'done', 'weight')])
        loss = api.q_1step_td_error(data, p['gamma'])
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'v1td':
        data = api.v_1step_td_data(t['v'], t['next_v'], t['reward'], t['done'], t['weight'])
        loss, per = api.v_1step_td_error(data, p['gamma'])
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'vntd':
        data = api.v_nstep_td_data(t['v'], t['next_n_v'], t['reward'], t['done'], t['weight'], t['value_gamma'])
        loss, per = api.v_nstep_td_error(data, p['gamma'], p['nstep'])
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'dntd':
        w = p.pop('weight_float', None)
        data = api.dist_nstep_td_data(t['dist'], t['next_n_dist'], t['act'], t['next_n_act'], t['reward'], t['done'],
                                      w if w is not None else t['weight'])
        if 'value_gamma' in t:
            p['value_gamma'] = t['value_gamma']
        loss, per = api.dist_nstep_td_error(data, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'td_lambda':
        loss = api.td_lambda_error(api.td_lambda_data(t['value'], t['reward'], t['weight']), **p)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'upgo':
        loss = api.upgo_loss(t['target_output'], t['rhos'], t['action'], t['rewards'], t['bootstrap_values'],
                             t['mask'])
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'vtrace':
        data = api.vtrace_data(t['target_output'], t['behaviour_output'], t['action'], t['value'], t['reward'],
                               t['weight'])
        loss = api.vtrace_error_discrete_action(data, **p)
        for k in ('policy_loss', 'value_loss', 'entropy_loss'):
            res['out_' + k] = _np(getattr(loss, k))
        _backward(op, list(loss), t, res)
        return res
    raise KeyError(op)


def run_oracle(orc, op, tensors, params):
    """Run a case through the flat-argument CPU oracle (``oracle/rl_oracle.py``)."""
    t = prepare(op, tensors, 'cpu')
    p = copy.copy(params)
    res = OrderedDict()
    if op == 'gae':
        adv = orc.gae(t['value'], t['next_value'], t['reward'], t['done'], t['traj_flag'], **p)
        res['out_adv'] = _np(adv)
        res['out_next_value_after'] = _np(t['next_value'])
        return res
    if op == 'ppo':
        out = orc.ppo_error(**t, **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss', 'kl_div'), out[:4]):
            res['out_' + k] = _np(v)
        res['out_approx_kl'] = np.float32(out[4])
        res['out_clipfrac'] = np.float32(out[5])
        _backward(op, list(out[:4]), t, res)
        return res
    if op in ('qntd', 'qntd_rescale'):
        fn = orc.q_nstep_td_error if op == 'qntd' else orc.q_nstep_td_error_with_rescale
        loss, per = fn(**t, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'ppoc':
        out = orc.ppo_error_continuous(**t, **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss', 'kl_div'), out[:4]):
            res['out_' + k] = _np(v)
        res['out_approx_kl'] = np.float32(out[4])
        res['out_clipfrac'] = np.float32(out[5])
        _backward(op, list(out[:4]), t, res)
        return res
    if op == 'vtc':
        out = orc.vtrace_error_continuous_action(**t, **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss'), out):
            res['out_' + k] = _np(v)
        _backward(op, list(out), t, res)
        return res
    if op == 'a2c':
        out = orc.a2c_error(**t)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss'), out):
            res['out_' + k] = _np(v)
        _backward(op, list(out), t, res)
        return res
    if op == 'retrace':
        res['out_q_retraces'] = _np(orc.compute_q_retraces(*t.values(), **p))
        return res
    if op == 'acer':
        actor, bias = orc.acer_policy_error(t['q_values'].detach(), t['q_retraces'], t['v_pred'], t['target_logit'], t['actions'],
                                            t['ratio'], p['c_clip_ratio'])
        critic = orc.acer_value_error(t['q_values'], t['q_retraces'], t['actions'])
        res['out_actor_loss'], res['out_bias_correction_loss'], res['out_critic_loss'] = _np(actor), _np(bias), _np(critic)
        w = torch.linspace(0.5, 1.5, actor.numel(), device=actor.device).reshape(actor.shape)
        ((actor * w).sum() + 0.3 * (bias * w.flip(0)).sum()).backward()
        (critic * w).sum().backward()
        res['grad_target_logit'], res['grad_q_values'] = _np(t['target_logit'].grad), _np(t['q_values'].grad)
        res['out_trust_region'] = _np(orc.acer_trust_region_update([t['actor_gradient']], t['target_logit'].detach(),
                                                                   t['avg_logit'], p['trust_region_value'])[0])
        return res
    if op == 'ppg':
        loss = orc.ppg_joint_error(**t, **p)
        res['out_auxiliary_loss'], res['out_behavioral_cloning_loss'] = _np(loss[0]), _np(loss[1])
        _backward(op, list(loss), t, res)
        return res
    if op == 'happoc':
        out = orc.happo_error_continuous(**t, **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss'), out[:3]):
            res['out_' + k] = _np(v)
        res['out_approx_kl'] = np.float32(out[3])
        res['out_clipfrac'] = np.float32(out[4])
        _backward(op, list(out[:3]), t, res)
        return res
    if op == 'happo':
        out = orc.happo_error(*[t[k] for k in HAPPO_FIELDS], **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss'), out[:3]):
            res['out_' + k] = _np(v)
        res['out_approx_kl'] = np.float32(out[3])
        res['out_clipfrac'] = np.float32(out[4])
        _backward(op, list(out[:3]), t, res)
        return res
    if op in ('qrdqn', 'iqn', 'fqf'):
        loss, per = getattr(orc, op + '_nstep_td_error')(*[t[k] for k in QUANTILE_FIELDS[op]], value_gamma=t.get('value_gamma'), **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'ppo_policy':
        out = orc.ppo_policy_error(**t, **p)
        for k, v in zip(('policy_loss', 'entropy_loss', 'kl_div'), out[:3]):
            res['out_' + k] = _np(v)
        res['out_approx_kl'] = np.float32(out[3])
        res['out_clipfrac'] = np.float32(out[4])
        _backward(op, list(out[:3]), t, res)
        return res
    if op == 'ppo_value':
        loss = orc.ppo_value_error(**t, **p)
        res['out_value_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'bdq':
        loss, per = orc.bdq_nstep_td_error(**t, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'qseq':
        loss, prio, per = orc.q_nstep_td_error_sequence(**t, **p)
        res['out_loss'] = _np(loss)
        res['out_priority'] = _np(prio)
        res['out_td_error'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'd1td':
        loss = orc.dist_1step_td_error(**t, **p)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'q1td':
        loss = orc.q_1step_td_error(**t, **p)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op in ('v1td', 'vntd'):
        fn = orc.v_1step_td_error if op == 'v1td' else orc.v_nstep_td_error
        loss, per = fn(**t, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'dntd':
        w = p.pop('weight_float', None)
        if w is not None:
            t['weight'] = w
        loss, per = orc.dist_nstep_td_error(**t, **p)
        res['out_loss'] = _np(loss)
        res['out_td_error_per_sample'] = _np(per)
        _backward(op, [loss], t, res)
        return res
    if op == 'td_lambda':
        loss = orc.td_lambda_error(**t, **p)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'upgo':
        loss = orc.upgo_loss(**t)
        res['out_loss'] = _np(loss)
        _backward(op, [loss], t, res)
        return res
    if op == 'vtrace':
        out = orc.vtrace_error_discrete_action(**t, **p)
        for k, v in zip(('policy_loss', 'value_loss', 'entropy_loss'), out):
            res['out_' + k] = _np(v)
        _backward(op, list(out), t, res)
        return res
    raise KeyError(op)


# outputs that are driven by integer / boolean decisions and must match bit-for-bit on the GPU as well
EXACT_KEYS = {'gae': ['out_adv', 'out_next_value_after'], 'retrace': ['out_q_retraces']}


def compare(res, ref, rtol=1e-5, atol=1e-5, exact=False):
    """Assert two result dicts agree. fp32 tolerance ``|a-b| <= atol + rtol*|b|`` (north star: 1e-5)."""
    assert set(res.keys()) == set(ref.keys()), (sorted(res.keys()), sorted(ref.keys()))
    for k in ref:
        a, b = np.asarray(res[k]), np.asarray(ref[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if exact:
            assert np.array_equal(a, b, equal_nan=True), (k, float(np.max(np.abs(a.astype(np.float64) - b))))
        else:
            # gradients of a mean over S samples scale like 1/S: make the absolute term relative to the tensor's scale
            # so that the check stays meaningful at S = 524288 (|grad| ~ 1e-6)
            scale = float(np.max(np.abs(b))) if (k.startswith('grad_') and b.size) else 1.0
            ok = np.allclose(a, b, rtol=rtol, atol=atol * scale, equal_nan=True)
            assert ok, (k, float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))), scale)
