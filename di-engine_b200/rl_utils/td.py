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
dist_1step_td_data = namedtuple(
    'dist_1step_td_data', ['dist', 'next_dist', 'act', 'next_act', 'reward', 'done', 'weight']
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


def _gamma_arg(gamma, B, dev, cum_reward):
    """python float -> (gamma, None); the NGU list of B 0-dim tensors (td.py:275-282) -> (0.0, (B,) device tensor)."""
    if isinstance(gamma, float):
        return gamma, None
    if isinstance(gamma, list):
        if cum_reward:
            raise TypeError("cum_reward with a list gamma is not defined by the reference (td.py:711-715)")
        gamma_ps = ops.f32c(torch.stack([torch.as_tensor(g) for g in gamma], dim=0).to(dev), 'gamma').reshape(-1)
        if gamma_ps.numel() != B:
            raise ValueError("list gamma must have B=%d entries" % B)
        return 0.0, gamma_ps
    raise TypeError("The type of gamma should be float or list")


def _qntd_rows(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, gamma, nstep, cum_reward,
               criterion, rescale, trans_fn=None, inv_trans_fn=None, group_mean=False):
    """The canonical row form every public signature below is brought to: q, next_n_q (S, G, N); action, next_n_action
    (S, G); reward (nstep, S) -- (S,) with cum_reward; done, weight (S,); value_gamma None | scalar | 1-element | (S,).
    Returns (loss, td) with td (S, G), or (S,) = mean over G with ``group_mean``; both carry gradient to ``q``."""
    dev = ops.compute_device(q, next_n_q)
    S, G, N = q.shape
    gamma_f, gamma_ps = _gamma_arg(gamma, S, dev, cum_reward)
    qd = ops.f32c(ops.to_device(q, dev), 'q')
    nq = ops.f32c(ops.to_device(next_n_q.detach(), dev), 'next_n_q')
    act = ops.i64c(ops.to_device(action, dev), q.shape[-1])
    nact = ops.i64c(ops.to_device(next_n_action, dev), q.shape[-1], 'next_n_action')
    r = ops.f32c(ops.to_device(reward.detach(), dev), 'reward')
    d = ops.f32c(ops.to_device(done.detach(), dev), 'done')
    if (nq.shape != qd.shape or act.numel() != S * G or nact.numel() != S * G or
            r.numel() != (S if cum_reward else nstep * S) or d.numel() != S):
        raise ValueError("q %s / next_n_q %s / action %s / next_n_action %s / reward %s / done %s do not match "
                         "S=%d, G=%d, nstep=%d" % (tuple(q.shape), tuple(next_n_q.shape), tuple(action.shape),
                                                   tuple(next_n_action.shape), tuple(reward.shape), tuple(done.shape),
                                                   S, G, nstep))
    w = None
    if weight is not None:
        w = ops.f32c(ops.to_device(weight.detach(), dev), 'weight')
        if w.numel() != S:
            w = w.expand(S).contiguous()
    vg, vg_stride = _value_gamma_arg(value_gamma, S, dev)
    custom_trans = rescale and (trans_fn is not value_transform or inv_trans_fn is not value_inv_transform)
    code = _criterion_code(criterion)
    gm = 1 if group_mean else 0

    def kernel(q_, nq_, act_, nact_, n_, resc, crit, param):
        return ops.QNStepTDFunction.apply(q_, nq_, act_, nact_, r, d, w, vg, vg_stride, gamma_ps, S, G, n_, int(nstep),
                                          float(gamma_f), 1 if cum_reward else 0, resc, 1e-2, crit, param, gm, 0, 0.0,
                                          False)

    if custom_trans or code is None:
        # user-supplied transforms / criterion are torch callables: the kernel produces the detached n-step target (of the
        # inverse-transformed next value), the callables run on the device around it
        wb = w.reshape(S, 1) if w is not None else 1.0
        if custom_trans:
            tq = inv_trans_fn(nq.gather(-1, nact.reshape(S, G, 1))).reshape(S, G, 1).contiguous()
            zero = torch.zeros(S, G, dtype=torch.int64, device=dev)
            _, _, target, _ = kernel(qd.detach().gather(-1, act.reshape(S, G, 1)).contiguous(), tq, zero, zero, 1, 0, 0,
                                     0.0)
            target = trans_fn(target.reshape(S, G)).detach()
        else:
            _, _, target, _ = kernel(qd.detach(), nq, act, nact, N, 1 if rescale else 0, 0, 0.0)
            target = target.reshape(S, G)
        per = criterion(qd.gather(-1, act.reshape(S, G, 1)).squeeze(-1), target)
        loss = (per * wb).mean()
        if group_mean:
            per = per.mean(-1)
        return loss, per
    loss, per, _, _ = kernel(qd, nq, act, nact, N, 1 if rescale else 0, code[0], code[1])
    return loss, (per if group_mean else per.reshape(S, G))


def _qntd(data, gamma, nstep, cum_reward, value_gamma, criterion, rescale, trans_fn=None, inv_trans_fn=None):
    """q_nstep_td_error / _with_rescale: the common (B, N) / (B,) call goes straight to the row form; every other shape the
    reference accepts (td.py:692-719) is first broadcast exactly as the reference's tensor expressions broadcast it."""
    q, next_n_q, action, next_n_action, reward, done, weight = data
    host_out = not q.is_cuda
    if not isinstance(gamma, (float, list)):
        raise TypeError("The type of gamma should be float or list")
    if not cum_reward:
        assert reward.shape[0] == nstep  # td.py:257
    B = q.shape[0]
    plain = (q.dim() == 2 and action.dim() == 1 and next_n_action.dim() == 1 and done.dim() == 1 and
             tuple(reward.shape) == ((B, ) if cum_reward else (nstep, B)) and
             (weight is None or weight.numel() in (1, B)))
    if plain:
        N = q.shape[1]
        loss, per = _qntd_rows(q.reshape(B, 1, N), next_n_q.reshape(B, 1, N), action.reshape(B, 1),
                               next_n_action.reshape(B, 1), reward, done, weight, value_gamma, gamma, nstep, cum_reward,
                               criterion, rescale, trans_fn, inv_trans_fn)
        per = per.reshape(B)
    else:
        loss, per = _qntd_general(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, gamma, nstep,
                                  cum_reward, criterion, rescale, trans_fn, inv_trans_fn)
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def _qntd_general(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, gamma, nstep, cum_reward,
                  criterion, rescale, trans_fn, inv_trans_fn):
    """Shapes beyond (B, N) / (B,): the reference evaluates q_nstep_td_error with broadcasting tensor expressions, so e.g.
    ``cum_reward=True`` with an (nstep, B) reward (its own test, tests/test_td.py:29-35) yields an (nstep, B) error, and the
    multi-agent branch (action (B, A, 1) against q (B, A, N), td.py:700-705) a (B, A) one.  Here the shape algebra of
    td.py:692-719 / :230-286 is replayed on the host (shapes only -- an incompatible combination raises the same
    broadcasting ``RuntimeError`` the reference raises), every operand is expanded to the resulting error shape and the
    rows go through the same kernel.  Expansion is a view + copy; the gradient of an expanded ``q`` is reduced by autograd."""
    bs = torch.broadcast_shapes
    weight_given = weight is not None
    w_shape = tuple(weight.shape) if weight_given else tuple(reward.shape)  # td.py:693 ones_like(reward)
    r_shape, d_shape = tuple(reward.shape), tuple(done.shape)
    vg_t = value_gamma if isinstance(value_gamma, torch.Tensor) else None
    vg_shape = tuple(vg_t.shape) if vg_t is not None else None
    act = action
    if action.dim() == 1 or action.dim() < q.dim():
        act = action.unsqueeze(-1)
    elif action.dim() > 1:  # the reference's multi-agent branch
        r_shape, w_shape, d_shape = r_shape + (1, ), w_shape + (1, ), d_shape + (1, )
        if vg_shape is not None:
            vg_shape = vg_shape + (1, )
    if act.dim() != q.dim() or tuple(act.shape[:-1]) != tuple(q.shape[:-1]) or act.shape[-1] != 1:
        raise RuntimeError("gather: action %s does not index q %s" % (tuple(action.shape), tuple(q.shape)))
    if tuple(next_n_action.shape) != tuple(next_n_q.shape[:-1]):
        raise RuntimeError("gather: next_n_action %s does not index next_n_q %s" %
                           (tuple(next_n_action.shape), tuple(next_n_q.shape)))
    qsa_shape, tq_shape = tuple(q.shape[:-1]), tuple(next_n_q.shape[:-1])

    def trail(shape, like):  # view_similar, td.py:222-224
        return tuple(shape) + (1, ) * (len(like) - len(shape))

    # each operand's shape as it enters the final (broadcasting) expression
    if cum_reward:
        rew_term = r_shape
        if vg_shape is None:
            val_term = bs(tq_shape, d_shape)
            vg_b, d_b = None, d_shape
        else:
            val_term = bs(vg_shape, tq_shape, d_shape)
            vg_b, d_b = vg_shape, d_shape
    else:
        rew_term = r_shape[1:]  # reward.mul(reward_factor).sum(0)
        if isinstance(gamma, list):
            g_b = trail((d_shape[0], ), r_shape[1:])  # reward_factor[nstep] of view_similar(reward_factor, reward)
            val_term = bs(g_b, tq_shape, d_shape)
            vg_b, d_b = None, d_shape
        elif value_gamma is None:
            val_term = bs(tq_shape, d_shape)
            vg_b, d_b = None, d_shape
        else:
            vg_b = trail(vg_shape if vg_shape is not None else tq_shape, tq_shape)  # np.isscalar -> full_like(next_value)
            d_b = trail(d_shape, tq_shape)
            val_term = bs(vg_b, tq_shape, d_b)
    td_shape = bs(qsa_shape, bs(rew_term, val_term))
    if bs(td_shape, w_shape) != td_shape:
        if weight_given:
            raise ValueError("weight %s does not broadcast to the td error shape %s" % (w_shape, tuple(td_shape)))
        # ones_like(reward) against a smaller error: the mean of the broadcast product is the plain mean
    S = 1
    for n_ in td_shape:
        S *= n_
    N = q.shape[-1]
    dev = ops.compute_device(q, next_n_q)

    def rows(x, shape, lead=()):  # expand `x` (viewed as `shape`) to lead + td_shape and flatten the td dims
        x = ops.to_device(x, dev)
        return x.reshape(tuple(lead) + tuple(shape)).expand(tuple(lead) + tuple(td_shape)).reshape(tuple(lead) + (S, ))

    q_rows = ops.to_device(q, dev).reshape(qsa_shape + (N, )).expand(tuple(td_shape) + (N, )).reshape(S, 1, N)
    nq_rows = ops.to_device(next_n_q, dev).expand(tuple(td_shape) + (N, )).reshape(S, 1, N)
    a_rows = rows(act.squeeze(-1), qsa_shape).reshape(S, 1)
    na_rows = rows(next_n_action, tq_shape).reshape(S, 1)
    r_rows = rows(reward, r_shape) if cum_reward else rows(reward, r_shape[1:], lead=(nstep, ))
    d_rows = rows(done.float(), d_b)
    w_rows = rows(weight, w_shape) if weight_given else None
    gam = gamma
    if isinstance(gamma, list):
        gam = list(rows(torch.stack([torch.as_tensor(g) for g in gamma], 0).float(), g_b))
    vg_rows = value_gamma
    if vg_t is not None and vg_t.numel() > 1:
        vg_rows = rows(vg_t, vg_b)
    loss, per = _qntd_rows(q_rows, nq_rows, a_rows, na_rows, r_rows, d_rows, w_rows, vg_rows, gam, nstep, cum_reward,
                           criterion, rescale, trans_fn, inv_trans_fn)
    return loss, per.reshape(tuple(td_shape))


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
    or (B,); gamma float, or the NGU list of B 0-dim tensors.  Every other shape combination the reference's
    broadcasting expressions accept is accepted too and gives the reference's result shape: ``cum_reward=True`` with an
    (nstep, B) reward (tests/test_td.py:29-35) -> (nstep, B) errors; the multi-agent branch q (B, A, N) with action
    (B, A, 1) (td.py:700-705) -> (B, A) errors; incompatible shapes raise the reference's broadcasting ``RuntimeError``.
    Returns ``(loss, td_error_per_sample)``, both differentiable w.r.t. ``q`` as in the reference (td.py:718-719).
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


def bdq_nstep_td_error(
        data: namedtuple,
        gamma: Union[float, list],
        nstep: int = 1,
        cum_reward: bool = False,
        value_gamma: Optional[torch.Tensor] = None,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
) -> torch.Tensor:
    """
    Multi-step TD error of the branching dueling Q-network (BDQ), drop-in for ding/rl_utils/td.py:722-789:
    q, next_n_q (B, D, N) -- D action branches of N bins; action, next_n_action (B, D); reward (nstep, B) -- (B,) with
    ``cum_reward``; done (B,); weight (B,) or None.  Every branch shares the sample's n-step reward; the per-sample error is
    the mean over the branches (td.py:788).  Returns ``(loss, td_error_per_sample (B,))``.  One launch on the q-n-step kernel
    (G = D rows per sample).
    """
    q, next_n_q, action, next_n_action, reward, done, weight = data
    host_out = not q.is_cuda
    if not isinstance(gamma, (float, list)):
        raise TypeError("The type of gamma should be float or list")
    if not cum_reward:
        assert reward.shape[0] == nstep  # td.py:257
    B = q.shape[0]
    if q.dim() != 3 or action.dim() != 2 or tuple(action.shape) != tuple(q.shape[:2]):
        raise RuntimeError("bdq_nstep_td_error: q %s / action %s must be (B, D, N) / (B, D)" %
                           (tuple(q.shape), tuple(action.shape)))
    if tuple(reward.shape) == ((B, ) if cum_reward else (nstep, B)):
        loss, per = _qntd_rows(q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, gamma, nstep,
                               cum_reward, criterion, False, group_mean=True)
    elif cum_reward and reward.dim() == 2 and reward.shape[1] == B:
        # (K, B) reward with cum_reward (tests/test_td.py:58-66): the reference broadcasts to a (K, B, D) error and averages
        # the branches -> (K, B); the K reward rows are K independent sample sets of the same q
        K, D, N = reward.shape[0], q.shape[1], q.shape[2]
        dev = ops.compute_device(q, next_n_q)

        def rep(x):
            x = ops.to_device(x, dev)
            return x.unsqueeze(0).expand((K, ) + tuple(x.shape)).reshape((K * B, ) + tuple(x.shape[1:]))

        vg = value_gamma
        if isinstance(vg, torch.Tensor) and vg.numel() > 1:
            vg = rep(vg)
        w = weight
        if w is not None:
            if tuple(w.shape) != (B, ):
                raise ValueError("weight %s does not broadcast to the td error shape %s" % (tuple(w.shape), (K, B)))
            w = rep(w)
        loss, per = _qntd_rows(rep(q), rep(next_n_q), rep(action), rep(next_n_action), ops.to_device(reward, dev).reshape(-1),
                               rep(done.float()), w, vg, gamma, nstep, True, criterion, False, group_mean=True)
        per = per.reshape(K, B)
    else:
        raise RuntimeError("bdq_nstep_td_error: reward %s does not match B=%d, nstep=%d" % (tuple(reward.shape), B, nstep))
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


q_nstep_td_seq_data = namedtuple(
    'q_nstep_td_seq_data', ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight']
)


def q_nstep_td_error_sequence(
        data: namedtuple,
        gamma: Union[float, list],
        nstep: int = 1,
        value_gamma: Optional[torch.Tensor] = None,
        rescale: bool = False,
        priority_mix: float = 0.9,
        criterion: torch.nn.modules = nn.MSELoss(reduction='none'),
):
    """
    The per-time-step loop of the recurrent Q-learners in ONE launch.  Not a reference function: it is exactly what
    ``R2D2Policy._forward_learn`` / NGU / R2D3 compute around the operator (ding/policy/r2d2.py:347-369, ngu.py:330-360)::

        for t in range(T):
            l, e = q_nstep_td_error[_with_rescale](q_nstep_td_data(q[t], next_n_q[t], action[t], next_n_action[t],
                                                   reward[t], done[t], weight[t]), gamma, nstep, value_gamma=value_gamma[t])
            loss.append(l); td_error.append(e.abs())
        loss = sum(loss) / (len(loss) + 1e-8)
        priority = mix * max_t(td_error) + (1 - mix) * sum_t(td_error) / (len(td_error) + 1e-8)

    Shapes: q, next_n_q (T, B, N); action, next_n_action (T, B); reward (T, nstep, B) (the policy's permuted layout,
    r2d2.py:343); done, weight (T, B) (weight may be None); value_gamma (T, B) or None; gamma float or the NGU list of B
    0-dim tensors.  Returns ``(loss, priority (B,), td_error (T, B))``: loss and td_error carry gradient to ``q``; the
    priority (the replay priority) is detached.
    """
    q, next_n_q, action, next_n_action, reward, done, weight = data
    host_out = not q.is_cuda
    if q.dim() != 3 or action.dim() != 2 or reward.dim() != 3:
        raise ValueError("q_nstep_td_error_sequence: expected q (T, B, N), action (T, B), reward (T, nstep, B); got %s / %s "
                         "/ %s" % (tuple(q.shape), tuple(action.shape), tuple(reward.shape)))
    T, B, N = q.shape
    assert reward.shape[1] == nstep  # td.py:257 for every step
    dev = ops.compute_device(q, next_n_q)
    gamma_f, gamma_ps = _gamma_arg(gamma, B, dev, False)
    code = _criterion_code(criterion)
    if code is None:
        raise NotImplementedError("q_nstep_td_error_sequence: criterion must be MSELoss / L1Loss / SmoothL1Loss / HuberLoss "
                                  "with reduction='none'")
    qd = ops.f32c(ops.to_device(q, dev), 'q')
    nq = ops.f32c(ops.to_device(next_n_q.detach(), dev), 'next_n_q')
    act = ops.i64c(ops.to_device(action, dev), q.shape[-1])
    nact = ops.i64c(ops.to_device(next_n_action, dev), q.shape[-1], 'next_n_action')
    r = ops.f32c(ops.to_device(reward.detach(), dev), 'reward')
    d = ops.f32c(ops.to_device(done.detach(), dev), 'done')
    S = T * B
    if (nq.shape != qd.shape or tuple(act.shape) != (T, B) or tuple(nact.shape) != (T, B) or
            tuple(r.shape) != (T, nstep, B) or d.numel() != S):
        raise ValueError("q_nstep_td_error_sequence: operand shapes do not match T=%d, B=%d, nstep=%d" % (T, B, nstep))
    w = None
    if weight is not None:
        w = ops.f32c(ops.to_device(weight.detach(), dev), 'weight')
        if w.numel() != S:
            w = w.expand(T, B).contiguous()
    vg, vg_stride = _value_gamma_arg(value_gamma, S, dev)
    loss, per, _, prio = ops.QNStepTDFunction.apply(
        qd, nq, act, nact, r, d, w, vg, vg_stride, gamma_ps, S, 1, N, int(nstep), float(gamma_f), 0,
        1 if rescale else 0, 1e-2, code[0], code[1], 0, T, float(priority_mix), True
    )
    per = per.reshape(T, B)
    if host_out:
        loss, prio, per = loss.cpu(), prio.cpu(), per.cpu()
    return loss, prio, per


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
    if weight is not None and not isinstance(weight, torch.Tensor):
        weight = ops.const_scalar(weight, dev).expand(S)  # a python-float weight (tests/test_td.py:389) scales every sample
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
    return _dntd(data, gamma, v_min, v_max, n_atom, nstep, value_gamma, check_positive=CHECK_DIST_POSITIVE)


def _dntd(data, gamma, v_min, v_max, n_atom, nstep, value_gamma, check_positive):
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
    a = ops.i64c(ops.to_device(act, dev), N, 'act')
    na = ops.i64c(ops.to_device(next_n_act, dev), N, 'next_n_act')
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
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if check_positive else None
    loss, per = ops.DistNStepTDFunction.apply(
        dd, nd, a, na, r, d, w, w_stride, vg, vg_stride, _support(v_min, v_max, n_atom, dev), B, A, N, int(n_atom),
        int(nstep), float(gamma), float(v_min), float(v_max), bad
    )
    if check_positive:
        assert bad.item() == 0, ("dist act", "non-positive probability in dist[batch_range, act]")  # td.py:513
    if host_out:
        loss, per = loss.cpu(), per.cpu()
    return loss, per


def dist_1step_td_error(
        data: namedtuple,
        gamma: float,
        v_min: float,
        v_max: float,
        n_atom: int,
) -> torch.Tensor:
    """
    1-step TD error of categorical (C51) distributional Q-learning, drop-in for ding/rl_utils/td.py:294-383: the
    ``nstep = 1`` case of ``dist_nstep_td_error`` (the reference pins that identity, tests/test_td.py:272-289) on the same
    kernel, returning the loss only.  dist, next_dist (B, N, n_atom) -- or (B, A, N, n_atom) with (B, A) actions; reward,
    done (B,); weight None (or a (B, 1) / 1-element tensor).  Unlike the n-step form the reference does not assert
    ``dist > 0`` here, so neither does this.
    """
    dist, next_dist, act, next_act, reward, done, weight = data
    assert len(reward.shape) == 1, reward.shape  # td.py:343
    if isinstance(weight, torch.Tensor) and weight.dim() == 2 and weight.shape[-1] == 1:
        weight = weight.squeeze(-1)
    loss, _ = _dntd(dist_nstep_td_data(dist, next_dist, act, next_act, reward.unsqueeze(0), done, weight), gamma, v_min,
                    v_max, n_atom, 1, None, check_positive=False)
    return loss


def _lambda_operand(x, like, dev, name):
    """gammas / lambda_ / done operand of the lambda-return: python scalar -> (None, scalar, False); tensor -> ((T,B), 0, rg)
    where rg says whether gradient has to flow back into it."""
    if x is None:
        return None, 0.0, False
    if not isinstance(x, torch.Tensor):
        return None, float(x), False
    rg = x.requires_grad and x.dtype.is_floating_point and torch.is_grad_enabled()
    t = ops.to_device(x if rg else x.detach(), dev)
    t = ops.f32c(t, name)
    if t.shape != like.shape:
        t = t.expand_as(like).contiguous()
    return t, 0.0, rg


def generalized_lambda_returns(
        bootstrap_values: torch.Tensor,
        rewards: torch.Tensor,
        gammas: float,
        lambda_: float,
        done: Optional[torch.Tensor] = None
) -> torch.Tensor:
    """
    Lambda-return G_t = r_t + (1-d_t)(gamma_t lambda_t G_{t+1} + gamma_t (1-lambda_t) V_{t+1}), drop-in for
    ding/rl_utils/td.py:1574-1651.  bootstrap_values (T+1, B); rewards (T, B) -- any trailing dims are accepted, e.g.
    Dreamer's (H, B, 1) (mbpolicy/utils.py:75); gammas / lambda_ floats or tensors shaped like rewards (bool lambdas as
    produced by UPGO are accepted); done None or shaped like rewards.  The forward values are bit-exact; like the
    reference function (plain torch arithmetic) the result is differentiable: gradients reach bootstrap_values, rewards
    and tensor gammas / lambda_ that require grad (MBSAC's actor loss back-propagates through it, mbpolicy/mbsac.py:137,153)
    through one transposed-scan launch.
    """
    dev = ops.compute_device(bootstrap_values, rewards)
    host_out = not bootstrap_values.is_cuda
    grad_on = torch.is_grad_enabled()
    rg_v = bootstrap_values.requires_grad and grad_on
    rg_r = rewards.requires_grad and grad_on
    v = ops.f32c(ops.to_device(bootstrap_values if rg_v else bootstrap_values.detach(), dev), 'bootstrap_values')
    r = ops.f32c(ops.to_device(rewards if rg_r else rewards.detach(), dev), 'rewards')
    if v.dim() < 1 or v.dim() != r.dim() or v.shape[0] != r.shape[0] + 1 or v.shape[1:] != r.shape[1:]:
        raise ValueError("expected bootstrap_values (T+1, B) and rewards (T, B), got %s / %s" %
                         (tuple(v.shape), tuple(r.shape)))
    out_shape = r.shape
    T = r.shape[0]
    v, r = v.reshape(T + 1, -1), r.reshape(T, -1)  # (T, B, ...) -> (T, B'): every trailing element is its own column

    def flat(x):
        return x.reshape(T, -1) if isinstance(x, torch.Tensor) and x.dim() == len(out_shape) else x

    gt, gs, rg_g = _lambda_operand(flat(gammas), r, dev, 'gammas')
    lt, ls, rg_l = _lambda_operand(flat(lambda_), r, dev, 'lambda_')
    dt = None
    if isinstance(done, torch.Tensor):
        dt, _, _ = _lambda_operand(flat(done.detach()), r, dev, 'done')
    elif done is not None:
        dt = torch.full_like(r, float(done))
    if rg_v or rg_r or rg_g or rg_l:
        ret = ops.LambdaReturnsFunction.apply(v, r, gt, lt, dt, gs, ls, False)
    else:
        ret = ops.lambda_returns_(v, r, gt, gs, lt, ls, dt, False)
    ret = ret.reshape(out_shape)
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
