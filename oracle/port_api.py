"""TEST / BENCH INFRASTRUCTURE ONLY -- the oracle port (oracle/rl_oracle.py) behind the reference's API names.

``bench.py``'s CPU arm times the unmodified reference whenever it is importable (the tree, or the byte-compiled archive
``oracle/_ref/ding_hotpath.zip``); only when neither exists does it fall back to this adapter (``cpu_baseline.kind = "port"``),
so that the two arms share one calling convention (``ding.rl_utils`` namedtuples and signatures)."""
from collections import namedtuple

from . import rl_oracle as _o

gae_data = namedtuple('gae_data', ['value', 'next_value', 'reward', 'done', 'traj_flag'])
ppo_data = namedtuple('ppo_data', ['logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight',
                                   'logit_pretrained'])
ppo_loss = namedtuple('ppo_loss', ['policy_loss', 'value_loss', 'entropy_loss', 'kl_div'])
ppo_info = namedtuple('ppo_info', ['approx_kl', 'clipfrac'])
q_nstep_td_data = namedtuple('q_nstep_td_data', ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight'])
dist_nstep_td_data = namedtuple('dist_1step_td_data', ['dist', 'next_n_dist', 'act', 'next_n_act', 'reward', 'done', 'weight'])
vtrace_data = namedtuple('vtrace_data', ['target_output', 'behaviour_output', 'action', 'value', 'reward', 'weight'])
vtrace_loss = namedtuple('vtrace_loss', ['policy_loss', 'value_loss', 'entropy_loss'])


def gae(data, gamma=0.99, lambda_=0.97):
    return _o.gae(*data, gamma=gamma, lambda_=lambda_)


def ppo_error(data, clip_ratio=0.2, use_value_clip=True, dual_clip=None, kl_type='k1'):
    out = _o.ppo_error(*data, clip_ratio=clip_ratio, use_value_clip=use_value_clip, dual_clip=dual_clip, kl_type=kl_type)
    return ppo_loss(*out[:4]), ppo_info(out[4], out[5])


def q_nstep_td_error(data, gamma, nstep=1, cum_reward=False, value_gamma=None):
    return _o.q_nstep_td_error(*data, gamma=gamma, nstep=nstep, cum_reward=cum_reward, value_gamma=value_gamma)


def dist_nstep_td_error(data, gamma, v_min, v_max, n_atom, nstep=1, value_gamma=None):
    return _o.dist_nstep_td_error(*data, gamma=gamma, v_min=v_min, v_max=v_max, n_atom=n_atom, nstep=nstep,
                                  value_gamma=value_gamma)


def vtrace_error_discrete_action(data, gamma=0.99, lambda_=0.95, rho_clip_ratio=1.0, c_clip_ratio=1.0,
                                 rho_pg_clip_ratio=1.0):
    return vtrace_loss(*_o.vtrace_error_discrete_action(*data, gamma=gamma, lambda_=lambda_, rho_clip_ratio=rho_clip_ratio,
                                                        c_clip_ratio=c_clip_ratio, rho_pg_clip_ratio=rho_pg_clip_ratio))
