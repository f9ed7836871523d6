"""n-step TD operators with the signatures of ding/rl_utils/td.py -- kernels in csrc/td.cu.

  q_nstep_td_error (td.py:649-719), q_nstep_td_error_with_rescale (:810-867), dist_nstep_td_error (:413-523),
  td_lambda_error (:1539-1571), generalized_lambda_returns (:1574-1605),
  and the 1-step / state-value siblings that reuse the q-n-step kernels:
  q_1step_td_error (:26-72), v_1step_td_error (:529-573), v_nstep_td_error (:579-617).
"""
from collections import namedtuple
from typing import Callable, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from .value_rescale import value_inv_transform, value_transform

q_nstep_td_data = namedtuple(
    'q_nstep_td_data', ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight']
)
# the reference's typename string is 'dist_1step_td_data' (td.py:386-388); kept for pickling / repr compatibility
dist_nstep_td_data = namedtuple(
    'dist_1step_td_data', ['dist', 'next_n_dist', 'act', 'next_n_act', 'reward', 'done', 'weight']
)
td_lambda_data = namedtuple('td_lambda_data', ['value', 'reward', 'weight'])
q_1step_td_data = namedtuple('q_1step_td_data', ['q', 'next_q', 'act', 'next_act', 'reward', 'done', 'weight'])
v_1step_td_data = namedtuple('v_1step_td_data', ['v', 'next_v', 'reward', 'done', 'weight'])
v_nstep_td_data = namedtuple('v_nstep_td_data', ['v', 'next_n_v', 'reward', 'done', 'weight', 'value_gamma'])

# The reference asserts ``dist[b, act] > 0`` on the host every call (td.py:513).  True keeps that behaviour (one
# 4-byte D2H read per call); False skips the read -- a non-positive entry then shows up as nan/inf in the loss.
CHECK_DIST_POSITIVE = True

_SUPPORT_CACHE = {}


def _data(args, kwargs):
    return args[0] if len(args) > 0 else kwargs['data']


def shape_fn_qntd(args, kwargs):
    """[T, B, N] cache key (td.py:632-645)."""
    d = _data(args, kwargs)
    return [d.reward.shape[0]] + list(d.q.shape)


def shape_fn_qntd_rescale(args, kwargs):
    """[T, B, N] cache key (td.py:791-804)."""
    d = _data(args, kwargs)
    return [d.reward.shape[0]] + list(d.q.shape)


def shape_fn_dntd(args, kwargs):
    """[T, B, N, n_atom] cache key (td.py:391-404)."""
    d = _data(args, kwargs)
    return [d.reward.shape[0]] + list(d.dist.shape)


def shape_fn_td_lambda(args, kwargs):
    """(T, B) cache key; the keyword form returns only T, exactly like the reference (td.py:1519-1529)."""
    if len(args) <= 0:
        return kwargs['data'].reward.shape[0]
    return args[0].reward.shape


def _criterion_code(criterion):
    """Map a torch criterion module to the fused kernel's enum; None when it has to run as a torch module."""
    if getattr(criterion, 'reduction', None) != 'none':
        return None
    if type(criterion) is nn.MSELoss:
        return 0, 0.0
    if type(criterion) is nn.L1Loss:
        return 1, 0.0
    if type(criterion) is nn.SmoothL1Loss:
        beta = float(criterion.beta)
        return (2, beta) if beta > 0 else (1, 0.0)
    if type(criterion) is nn.HuberLoss:
        return 3, float(criterion.delta)
    return None


def _value_gamma_arg(value_gamma, B, dev):
    """-> (tensor or None, stride): None | python scalar | 0-dim / 1-element tensor (stride 0) | (B,) tensor (stride 1)."""
    if value_gamma is None:
        return None, 0
    if not isinstance(value_gamma, torch.Tensor):
        return ops.const_scalar(value_gamma, dev), 0
    vg = ops.f32c(ops.to_device(value_gamma.detach(), dev), 'value_gamma')
    if vg.numel() == 1:
        return vg.reshape(1), 0
    if vg.numel() != B:
        raise ValueError("value_gamma must have 1 or B=%d elements, got %s" % (B, tuple(value_gamma.shape)))
    return vg.reshape(B), 1


def _qntd(data, gamma, nstep, cum_reward, value_gamma, criterion, rescale, trans_fn=None, inv_trans_fn=None):
    q, next_n_q, action, next_n_action, reward, done, weight = data
    if action.dim() != 1 or q.dim() != 2:
        raise NotImplementedError(
            "di_engine_b200.q_nstep_td_error: only q (B, N) with action (B,) is implemented on the B200 path "
            "(got q %s, action %s)" % (tuple(q.shape), tuple(action.shape))
        )
    dev = ops.compute_device(q, next_n_q)
    host_out = not q.is_cuda
    B, N = q.shape
    gamma_ps = None
    if isinstance(gamma, float):
        gamma_f = gamma
    elif isinstance(gamma, list):  # NGU: one 0-dim tensor per sample (td.py:275-282)
        if cum_reward:
            raise TypeError("cum_reward with a list gamma is not defined by the reference (td.py:711-715)")
        gamma_ps = ops.f32c(torch.stack([torch.as_tensor(g) for g in gamma], dim=0).to(dev), 'gamma').reshape(-1)
        if gamma_ps.numel() != B:
            raise ValueError("list gamma must have B=%d entries" % B)
        gamma_f = 0.0
    else:
        raise TypeError("The type of gamma should be float or list")
    if not cum_reward:
        assert reward.shape[0] == nstep  # td.py:257
    qd = ops.f32c(ops.to_device(q, dev), 'q')
    nq = ops.f32c(ops.to_device(next_n_q.detach(), dev), 'next_n_q')
    act = ops.i64c(ops.to_device(action, dev))
    nact = ops.i64c(ops.to_device(next_n_action, dev))
    r = ops.f32c(ops.to_device(reward.detach(), dev), 'reward')
    d = ops.f32c(ops.to_device(done.detach(), dev), 'done')
    if r.numel() != (B if cum_reward else nstep * B) or d.numel() != B:
        raise ValueError("reward %s / done %s do not match B=%d, nstep=%d" % (tuple(reward.shape), tuple(done.shape), B,
                                                                            nstep))
    w = None
    if weight is not None:
        w = ops.f32c(ops.to_device(weight.detach(), dev), 'weight')
        if w.numel() != B:
            w = w.expand(B).contiguous()
    vg, vg_stride = _value_gamma_arg(value_gamma, B, dev)
    custom_trans = rescale and (trans_fn is not value_transform or inv_trans_fn is not value_inv_transform)
    code = _criterion_code(criterion)
    if custom_trans:
        # user-supplied transforms are torch callables: evaluate them on the device around the fused target kernel
        rows = torch.arange(B, device=dev)
        tq = inv_trans_fn(nq[rows, nact]).reshape(B, 1).contiguous()
        zero = torch.zeros(B, dtype=torch.int64, device=dev)
        _, _, target = ops.QNStepTDFunction.apply(
            qd.detach()[rows, act].reshape(B, 1).contiguous(), tq, zero, zero, r, d, w, vg, vg_stride, gamma_ps,
            int(nstep), float(gamma_f), 0, 0, 0.0, 0, 0.0
        )
        per = criterion(qd[rows, act], trans_fn(target).detach())
        loss = (per * (w if w is not None else 1.0)).mean()
    elif code is None:
        # arbitrary criterion module: the kernel produces the detached n-step target, the module runs on the device
        _, _, target = ops.QNStepTDFunction.apply(
            qd.detach(), nq, act, nact, r, d, w, vg, vg_stride, gamma_ps, int(nstep), float(gamma_f),
            1 if cum_reward else 0, 1 if rescale else 0, 1e-2, 0, 0.0
        )
        per = criterion(qd.gather(-1, act.unsqueeze(-1)).squeeze(-1), target)
        loss = (per * (w if w is not None else 1.0)).mean()
    else:
        loss, per, _ = ops.QNStepTDFunction.apply(
            qd, nq, act, nact, r, d, w, vg, vg_stride, gamma_ps, int(nstep), float(gamma_f), 1 if cum_reward else 0,
            1 if rescale else 0, 1e-2, code[0], code[1]
        )
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def q_nstep_td_error(
        data: namedtuple,
        gamma: Union[float, list],
        nstep: int = 1,
        cum_reward: bool = False,
        value_gamma: Optional[torch.Tensor] = None,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
) -> torch.Tensor:
    """
    Multi-step TD error for Q-learning, drop-in for ding/rl_utils/td.py:649-719 (n-step return :230-286).

    Shapes: q, next_n_q (B, N); action, next_n_action (B,) int64; reward (nstep, B) -- (B,) with ``cum_reward``;
    done (B,) (a float multiplier, as the reference's tests pass it); weight (B,) or None; value_gamma None, scalar
    or (B,); gamma float, or the NGU list of B 0-dim tensors.  Returns ``(loss, td_error_per_sample)``; the loss is
    differentiable w.r.t. ``q`` (``td_error_per_sample`` is returned detached -- every caller in ding.policy only
    uses it for priorities).
    """
    return _qntd(data, gamma, nstep, cum_reward, value_gamma, criterion, rescale=False)


def q_nstep_td_error_with_rescale(
    data: namedtuple,
    gamma: Union[float, list],
    nstep: int = 1,
    value_gamma: Optional[torch.Tensor] = None,
    criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
    trans_fn: Callable = value_transform,
    inv_trans_fn: Callable = value_inv_transform,
) -> torch.Tensor:
    """
    Multi-step TD error with value rescaling h / h^-1, drop-in for ding/rl_utils/td.py:810-867.
    With the default transforms everything (h^-1, n-step return, h, criterion) runs in one kernel.
    """
    assert len(data.action.shape) == 1, data.action.shape  # td.py:854
    return _qntd(data, gamma, nstep, False, value_gamma, criterion, True, trans_fn, inv_trans_fn)


def q_1step_td_error(
        data: namedtuple,
        gamma: float,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
) -> torch.Tensor:
    """
    1-step TD error for Q-learning, drop-in for ding/rl_utils/td.py:26-72: ``target = gamma*(1-done)*next_q[next_act] +
    reward`` -- the n = 1 case of ``q_nstep_td_error`` (the reference pins that identity, tests/test_td.py:113-126), on the
    same kernel.  q, next_q (B, N); act, next_act (B,) int64; reward, done (B,); weight (B,) or None.  Returns the loss only.
    """
    q, next_q, act, next_act, reward, done, weight = data
    assert len(act.shape) == 1, act.shape        # td.py:64
    assert len(reward.shape) == 1, reward.shape  # td.py:65
    loss, _ = _qntd(q_nstep_td_data(q, next_q, act, next_act, reward.unsqueeze(0), done, weight), gamma, 1, False, None,
                    criterion, False)
    return loss


def _as_columns(v, next_v, reward, done, weight, value_gamma=None, nstep=None):
    """State values as a one-action Q table: (S, 1) q / next_q with action 0, per-sample tensors flattened to S = v.numel().
    ``reward``/``done``/``value_gamma`` of shape (B,) against v (B, K) are repeated along K like the reference's
    ``unsqueeze(1)`` broadcast (td.py:563-567)."""
    dev = ops.compute_device(v, next_v)
    S = v.numel()

    def per_sample(x, lead=0):
        if x is None:
            return None
        x = ops.to_device(x.detach() if isinstance(x, torch.Tensor) else x, dev)
        if x.dim() - lead < v.dim():  # (B,) against (B, K): broadcast over the trailing dims of v
            x = x.reshape(tuple(x.shape) + (1, ) * (v.dim() - (x.dim() - lead))).expand(tuple(x.shape[:lead]) + tuple(v.shape))
        return x.reshape(tuple(x.shape[:lead]) + (S, ))

    zero = torch.zeros(S, dtype=torch.int64, device=dev)
    r = per_sample(reward, lead=0 if nstep is None else 1)
    if nstep is None:
        r = r.unsqueeze(0)
    d = per_sample(done)
    if d is None:
        d = torch.zeros(S, dtype=torch.float32, device=dev)
    w = per_sample(weight)
    vg = value_gamma
    if isinstance(vg, torch.Tensor) and vg.numel() > 1:
        vg = per_sample(vg)
    return q_nstep_td_data(ops.to_device(v, dev).reshape(S, 1), ops.to_device(next_v.detach(), dev).reshape(S, 1), zero,
                           zero, r, d, w), vg


def v_1step_td_error(
        data: namedtuple,
        gamma: float,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
) -> torch.Tensor:
    """
    1-step TD error for a state-value (or critic) head, drop-in for ding/rl_utils/td.py:529-573 -- the critic loss of
    DDPG / TD3 / SAC / ...: ``target = gamma*(1-done)*next_v + reward``.  v, next_v (B,) or (B, K) with reward, done (B,)
    broadcast over K; done and weight may be None.  Returns ``(loss, td_error_per_sample)`` (the latter detached and
    shaped like ``v``).  Runs on the q-n-step kernel with one action column.
    """
    v, next_v, reward, done, weight = data
    host_out = not v.is_cuda
    qd, _ = _as_columns(v, next_v, reward, done, weight)
    loss, per = _qntd(qd, gamma, 1, False, None, criterion, False)
    per = per.reshape(v.shape)
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def v_nstep_td_error(
        data: namedtuple,
        gamma: float,
        nstep: int = 1,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
) -> torch.Tensor:
    """
    n-step TD error for a state-value head, drop-in for ding/rl_utils/td.py:579-617: ``target = nstep_return(reward,
    next_n_v, done, gamma, nstep, value_gamma)``.  v, next_n_v, done (B,); reward (nstep, B); weight, value_gamma (B,) or
    None.  Returns ``(loss, td_error_per_sample)``.  Runs on the q-n-step kernel with one action column.
    """
    v, next_n_v, reward, done, weight, value_gamma = data
    host_out = not v.is_cuda
    qd, vg = _as_columns(v, next_n_v, reward, done, weight, value_gamma, nstep=nstep)
    loss, per = _qntd(qd, gamma, nstep, False, vg, criterion, False)
    per = per.reshape(v.shape)
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def _support(v_min, v_max, n_atom, dev):
    key = (float(v_min), float(v_max), int(n_atom), dev.index)
    s = _SUPPORT_CACHE.get(key)
    if s is None:
        # built on the host then moved, exactly like td.py:457 -- CPU and CUDA linspace differ in the last bit
        s = torch.linspace(v_min, v_max, n_atom).to(dev)
        _SUPPORT_CACHE[key] = s
    return s


def dist_nstep_td_error(
        data: namedtuple,
        gamma: float,
        v_min: float,
        v_max: float,
        n_atom: int,
        nstep: int = 1,
        value_gamma: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """
    Multi-step TD error of categorical (C51) distributional Q-learning, drop-in for ding/rl_utils/td.py:413-523.

    Shapes: dist, next_n_dist (B, N, n_atom) probabilities -- or (B, A, N, n_atom) with (B, A) actions
    (td.py:470-489); act, next_n_act (B,) int64; reward (nstep, B); done (B,); weight None, python float, 1-element
    or per-row tensor; value_gamma None, float, 0-dim or (B,) tensor.
    Returns ``(loss, td_error_per_sample)`` -- weighted mean loss, UNWEIGHTED per-sample error (td.py:519-521).
    """
    dist, next_n_dist, act, next_n_act, reward, done, weight = data
    dev = ops.compute_device(dist, next_n_dist)
    host_out = not dist.is_cuda
    if act.dim() == 1:
        B, A = act.shape[0], 1
        N = dist.shape[1]
    else:
        B, A = act.shape
        N = dist.shape[2]
    R = B * A
    if dist.shape[-1] != n_atom or dist.numel() != R * N * n_atom or next_n_dist.numel() != dist.numel():
        raise ValueError("dist %s does not match act %s / n_atom=%d" % (tuple(dist.shape), tuple(act.shape), n_atom))
    assert reward.shape[0] == nstep and reward.numel() == nstep * B, reward.shape
    dd = ops.f32c(ops.to_device(dist, dev), 'dist')
    nd = ops.f32c(ops.to_device(next_n_dist.detach(), dev), 'next_n_dist')
    a = ops.i64c(ops.to_device(act, dev))
    na = ops.i64c(ops.to_device(next_n_act, dev))
    r = ops.f32c(ops.to_device(reward.detach(), dev), 'reward')
    d = ops.f32c(ops.to_device(done.detach(), dev), 'done')
    if weight is None:
        w, w_stride = None, 0
    elif not isinstance(weight, torch.Tensor):
        w, w_stride = ops.const_scalar(weight, dev), 0
    else:
        w = ops.f32c(ops.to_device(weight.detach(), dev), 'weight')
        if w.numel() == 1:
            w, w_stride = w.reshape(1), 0
        elif w.numel() == R:
            w, w_stride = w.reshape(R), 1
        else:
            raise ValueError("weight must have 1 or %d elements, got %s" % (R, tuple(weight.shape)))
    vg, vg_stride = _value_gamma_arg(value_gamma, B, dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    loss, per = ops.DistNStepTDFunction.apply(
        dd, nd, a, na, r, d, w, w_stride, vg, vg_stride, _support(v_min, v_max, n_atom, dev), B, A, N, int(n_atom),
        int(nstep), float(gamma), float(v_min), float(v_max), bad
    )
    if CHECK_DIST_POSITIVE:
        assert bad.item() == 0, ("dist act", "non-positive probability in dist[batch_range, act]")  # td.py:513
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def _tb_operand(x, like, dev, name):
    """gammas / lambda_ / done operand of the lambda-return: python scalar -> (None, scalar), tensor -> ((T,B), 0)."""
    if x is None:
        return None, 0.0
    if not isinstance(x, torch.Tensor):
        return None, float(x)
    t = ops.f32c(ops.to_device(x.detach(), dev), name)
    if t.shape != like.shape:
        t = t.expand_as(like).contiguous()
    return t, 0.0


def generalized_lambda_returns(
        bootstrap_values: torch.Tensor,
        rewards: torch.Tensor,
        gammas: float,
        lambda_: float,
        done: Optional[torch.Tensor] = None
) -> torch.Tensor:
    """
    Lambda-return G_t = r_t + (1-d_t)(gamma_t lambda_t G_{t+1} + gamma_t (1-lambda_t) V_{t+1}), drop-in for
    ding/rl_utils/td.py:1574-1651.  bootstrap_values (T+1, B); rewards (T, B); gammas / lambda_ floats or (T, B)
    tensors (bool lambdas as produced by UPGO are accepted); done None or (T, B).  Bit-exact, no autograd graph.
    """
    dev = ops.compute_device(bootstrap_values, rewards)
    host_out = not bootstrap_values.is_cuda
    v = ops.f32c(ops.to_device(bootstrap_values.detach(), dev), 'bootstrap_values')
    r = ops.f32c(ops.to_device(rewards.detach(), dev), 'rewards')
    if v.dim() != 2 or r.dim() != 2 or v.shape[0] != r.shape[0] + 1 or v.shape[1] != r.shape[1]:
        raise ValueError("expected bootstrap_values (T+1, B) and rewards (T, B), got %s / %s" %
                         (tuple(v.shape), tuple(r.shape)))
    gt, gs = _tb_operand(gammas, r, dev, 'gammas')
    lt, ls = _tb_operand(lambda_, r, dev, 'lambda_')
    dt, _ = _tb_operand(done, r, dev, 'done') if isinstance(done, torch.Tensor) else (None, 0.0)
    if done is not None and not isinstance(done, torch.Tensor):
        dt = torch.full_like(r, float(done))
    ret = ops.lambda_returns_(v, r, gt, gs, lt, ls, dt, False)
    return ret.cpu() if host_out else ret


def td_lambda_error(data: namedtuple, gamma: float = 0.9, lambda_: float = 0.8) -> torch.Tensor:
    """
    TD(lambda) loss 0.5 * mean(w * (G^lambda - V_{:-1})^2), drop-in for ding/rl_utils/td.py:1539-1571.
    value (T+1, B) (gradient reaches value[:-1]); reward (T, B); weight None or broadcastable to (T, B).
    Scan and loss head run in one kernel.
    """
    value, reward, weight = data
    dev = ops.compute_device(value, reward)
    host_out = not value.is_cuda
    v = ops.f32c(ops.to_device(value, dev), 'value')
    r = ops.f32c(ops.to_device(reward.detach(), dev), 'reward')
    if v.dim() != 2 or r.dim() != 2 or v.shape[0] != r.shape[0] + 1 or v.shape[1] != r.shape[1]:
        raise ValueError("expected value (T+1, B) and reward (T, B), got %s / %s" % (tuple(v.shape), tuple(r.shape)))
    w = None
    if weight is not None:
        w = ops.f32c(ops.to_device(weight.detach(), dev), 'weight')
        if w.shape != r.shape:
            w = w.expand_as(r).contiguous()
    loss = ops.td_lambda_(v, r, w, gamma, lambda_)
    return loss.cpu() if host_out else loss
