"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DI-engine learner hot path (``ding.rl_utils``).

This module is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it. The product package (``di-engine_b200``) never
does; it fails loudly when its CUDA library is missing.

Every function restates one reference operator with plain fp32 torch-CPU primitives, in the reference's
evaluation order (so that on CPU it is bit-identical to the reference, which
``tests/test_oracle.py::test_oracle_bit_exact_vs_live_reference`` asserts for every seeded case whenever the reference is
importable -- the tree here, the byte-compiled archive ``oracle/_ref/ding_hotpath.zip`` on the GPU box -- and which the
committed fixtures in ``tests/golden/`` pin as well).  Parity status: **pinned** -- the reference has no golden vectors of
its own for this path, so the pins are outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py``.  Exception, marked where it occurs: the collector-side restatements at the end of the file
(``default_collate_flat``, ``default_preprocess_learn``, ``adder_get_nstep_return_data``) are **parity unpinned** -- their
reference modules import packages that are not installed here, so the live code cannot run next to them.

Citations are ``file:line`` relative to ``/root/reference/ding/rl_utils/``.
Inputs are torch tensors; tensors that need gradients must have ``requires_grad`` set by the caller and the
returned losses are differentiable through autograd (the oracle for the CUDA backward kernels).
"""
import math
from typing import Optional, Sequence, Union

import numpy as np
import torch

_F32_MIN = torch.finfo(torch.float32).min


# --------------------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------------------
def _trailing_ones(x: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """td.py:222-224 ``view_similar``: append singleton dims so ``x`` broadcasts against ``like``."""
    return x.reshape(tuple(x.shape) + (1, ) * (like.dim() - x.dim()))


def _log_softmax_rows(logit: torch.Tensor) -> torch.Tensor:
    """``Categorical(logits=x).logits`` == ``x - logsumexp(x, -1, keepdim=True)`` (torch.distributions)."""
    return logit - logit.logsumexp(dim=-1, keepdim=True)


def _chosen(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return table.gather(-1, idx.long().unsqueeze(-1)).squeeze(-1)


def _row_entropy(logp: torch.Tensor) -> torch.Tensor:
    """``Categorical.entropy``: -(clamp(logp, min=finfo.min) * softmax).sum(-1)."""
    probs = torch.softmax(logp, dim=-1)
    return -(logp.clamp(min=_F32_MIN) * probs).sum(-1)


# --------------------------------------------------------------------------------------------------------------
# gae.py:25-70
# --------------------------------------------------------------------------------------------------------------
def gae(value, next_value, reward, done=None, traj_flag=None, gamma: float = 0.99, lambda_: float = 0.97):
    """Generalised advantage estimate, reverse recurrence along dim 0.

    gae.py:51-59 defaults/broadcast; :61 masks ``next_value`` IN PLACE; :62 delta; :63 trace factor; :67-69 the
    loop ``A_t = delta_t + factor_t * A_{t+1}``.
    """
    if done is None:
        done = torch.zeros_like(reward)
    if traj_flag is None:
        traj_flag = done
    done = done.float()
    traj_flag = traj_flag.float()
    if value.dim() == reward.dim() + 1:  # (T,B,A) values with (T,B) rewards
        reward, done, traj_flag = reward.unsqueeze(-1), done.unsqueeze(-1), traj_flag.unsqueeze(-1)
    next_value *= (1 - done)  # caller-visible mutation, as in the reference
    delta = reward + gamma * next_value - value
    factor = gamma * lambda_ * (1 - traj_flag)
    adv = torch.zeros_like(value)
    carry = torch.zeros_like(value[0])
    for t in range(reward.shape[0] - 1, -1, -1):
        carry = delta[t] + factor[t] * carry
        adv[t] = carry
    return adv


# --------------------------------------------------------------------------------------------------------------
# the batch-level pieces PPOPolicy._forward_learn wraps around gae / ppo_error (ding/policy/ppo.py:274-306)
# --------------------------------------------------------------------------------------------------------------
def ppo_policy_gae_returns(value, next_value, reward, done, traj_flag, gamma, lambda_, std=None):
    """policy/ppo.py:274-297 (recompute_adv branch).  ``std`` = RunningMeanStd.std (a python / numpy float) or None when
    value_norm is off.  Returns (adv, value, return_, unnormalized_returns, (batch_mean, batch_var, batch_count)) where the
    statistics are what RunningMeanStd.update (utils/default_helper.py:547-567) derives from the array -- over ALL elements."""
    import numpy as np
    value, next_value = value.clone(), next_value.clone()
    if std is not None:
        value *= std  # :277
        next_value *= std  # :278
    adv = gae(value, next_value, reward, done, traj_flag, gamma, lambda_)  # :280-282
    unnorm = value + adv  # :284
    if std is not None:
        v, ret = value / std, unnorm / std  # :287-288
    else:
        v, ret = value, unnorm  # :291-292
    x = unnorm.numpy().reshape(-1)
    return adv, v, ret, unnorm, (float(np.mean(x)), float(np.var(x)), float(x.shape[0]))


def impala_reshape_data(values, rewards, done):
    """policy/impala.py:316-322 (on a copy: the reference multiplies ``values[1:]`` in place)."""
    weights_ = 1 - done.float()
    weights = torch.ones_like(rewards)
    values = torch.cat([values[:1], values[1:] * weights_], 0)
    weights = torch.cat([weights[:1], weights_[:-1]], 0)
    rewards = rewards * weights
    return values, rewards, weights


def normalize_advantage(adv):
    """policy/ppo.py:304-306: ``(adv - adv.mean()) / (adv.std() + 1e-8)`` (torch.std: unbiased)."""
    return (adv - adv.mean()) / (adv.std() + 1e-8)


# --------------------------------------------------------------------------------------------------------------
# ppo.py:77-275
# --------------------------------------------------------------------------------------------------------------
def ppo_error(
        logit_new,
        logit_old,
        action,
        value_new,
        value_old,
        adv,
        return_,
        weight=None,
        logit_pretrained=None,
        clip_ratio: float = 0.2,
        use_value_clip: bool = True,
        dual_clip: Optional[float] = None,
        kl_type: str = 'k1'
):
    """Returns ``(policy_loss, value_loss, entropy_loss, kl_div, approx_kl, clipfrac)``.

    The four losses are 0-dim tensors (differentiable), the two infos python floats (ppo.py:217-220).
    """
    assert dual_clip is None or dual_clip > 1.0, "dual_clip value must be greater than 1.0, but get value: {}".format(
        dual_clip
    )  # ppo.py:129
    w_pol = torch.ones_like(adv) if weight is None else weight  # ppo.py:190-191
    lp_new_all = _log_softmax_rows(logit_new)
    lp_old_all = _log_softmax_rows(logit_old)
    lp_new = _chosen(lp_new_all, action)  # ppo.py:194
    lp_old = _chosen(lp_old_all, action)  # ppo.py:195
    ent = _row_entropy(lp_new_all)  # ppo.py:198
    if ent.shape != w_pol.shape:  # ppo.py:199-200 multi-agent: average over the agent axis
        ent = ent.mean(dim=1)
    entropy_loss = (ent * w_pol).mean()  # ppo.py:201
    ratio = torch.exp(lp_new - lp_old)  # ppo.py:205
    if ratio.shape != adv.shape:
        ratio = ratio.mean(dim=1)  # ppo.py:206-207
    s1 = ratio * adv
    s2 = ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv
    if dual_clip is not None:  # ppo.py:210-214
        inner = torch.min(s1, s2)
        outer = torch.max(inner, dual_clip * adv)
        policy_loss = -(torch.where(adv < 0, outer, inner) * w_pol).mean()
    else:  # ppo.py:216
        policy_loss = (-torch.min(s1, s2) * w_pol).mean()
    with torch.no_grad():  # ppo.py:217-220
        approx_kl = (lp_old - lp_new).mean().item()
        clipped = ratio.gt(1 + clip_ratio) | ratio.lt(1 - clip_ratio)
        clipfrac = clipped.float().mean().item()
    if logit_pretrained is not None:  # ppo.py:222-226 + calculate_kl_div :30-54
        lp_pre = _chosen(_log_softmax_rows(logit_pretrained), action)
        log_ratio = lp_new - lp_pre
        if kl_type == 'k1':
            kl_div = log_ratio.mean()
        elif kl_type == 'k2':
            kl_div = (log_ratio ** 2 / 2).mean()
        elif kl_type == 'k3':
            kl_div = (torch.exp(-log_ratio) - 1 + log_ratio).mean()
        else:
            raise ValueError(f"Unknown kl_type: {kl_type}")
    else:
        kl_div = torch.tensor(0., dtype=policy_loss.dtype, device=policy_loss.device)  # ppo.py:228

    w_val = torch.ones_like(value_old) if weight is None else weight  # ppo.py:264-265
    if use_value_clip:  # ppo.py:267-272
        v_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        e1 = (return_ - value_new).pow(2)
        e2 = (return_ - v_clip).pow(2)
        value_loss = 0.5 * (torch.max(e1, e2) * w_val).mean()
    else:  # ppo.py:274
        value_loss = 0.5 * ((return_ - value_new).pow(2) * w_val).mean()
    return policy_loss, value_loss, entropy_loss, kl_div, approx_kl, clipfrac


# --------------------------------------------------------------------------------------------------------------
# td.py:230-286 nstep_return ; value_rescale.py:4-34
# --------------------------------------------------------------------------------------------------------------
def ppo_error_continuous(mu_new, sigma_new, mu_old, sigma_old, action, value_new, value_old, adv, return_, weight=None,
                         mu_pretrained=None, sigma_pretrained=None, clip_ratio: float = 0.2, use_value_clip: bool = True,
                         dual_clip: Optional[float] = None, kl_type: str = 'k1'):
    """ppo.py:278-374 with the Independent(Normal) log-prob / entropy written out (torch.distributions.Normal.log_prob:
    ``-((x-mu)^2)/(2 var) - log(sigma) - log(sqrt(2 pi))``; entropy ``0.5 + 0.5 log(2 pi) + log(sigma)``; Independent sums the
    last dim).  Returns (policy, value, entropy, kl, approx_kl, clipfrac)."""
    assert dual_clip is None or dual_clip > 1.0
    if weight is None:
        weight = torch.ones_like(adv)

    def logp(mu, sigma):
        var = sigma ** 2
        return (-((action - mu) ** 2) / (2 * var) - sigma.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)

    if mu_old.dim() == 1:  # ppo.py:336-337
        mu_old, sigma_old = mu_old.unsqueeze(-1), sigma_old.unsqueeze(-1)
    logp_new, logp_old = logp(mu_new, sigma_new), logp(mu_old, sigma_old)
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma_new)).sum(-1)
    entropy_loss = (entropy * weight).mean()
    ratio = torch.exp(logp_new - logp_old)
    surr1 = ratio * adv
    surr2 = ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv
    if dual_clip is not None:
        policy_loss = (-torch.max(torch.min(surr1, surr2), dual_clip * adv) * weight).mean()
    else:
        policy_loss = (-torch.min(surr1, surr2) * weight).mean()
    with torch.no_grad():
        approx_kl = (logp_old - logp_new).mean().item()
        clipped = ratio.gt(1 + clip_ratio) | ratio.lt(1 - clip_ratio)
        clipfrac = torch.as_tensor(clipped).float().mean().item()
    if use_value_clip:
        value_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        value_loss = 0.5 * (torch.max((return_ - value_new).pow(2), (return_ - value_clip).pow(2)) * weight).mean()
    else:
        value_loss = 0.5 * ((return_ - value_new).pow(2) * weight).mean()
    if mu_pretrained is not None:
        log_ratio = logp_new - logp(mu_pretrained, sigma_pretrained)
        if kl_type == 'k1':
            kl_div = log_ratio.mean()
        elif kl_type == 'k2':
            kl_div = (log_ratio ** 2 / 2).mean()
        elif kl_type == 'k3':
            kl_div = (torch.exp(-log_ratio) - 1 + log_ratio).mean()
        else:
            raise ValueError(f"Unknown kl_type: {kl_type}")
    else:
        kl_div = torch.tensor(0.)
    return policy_loss, value_loss, entropy_loss, kl_div, approx_kl, clipfrac


def a2c_error(logit, action, value, adv, return_, weight=None):
    """a2c.py:10-44."""
    if weight is None:
        weight = torch.ones_like(value)
    logp_all = _log_softmax_rows(logit)
    logp = _chosen(logp_all, action)
    entropy_loss = (_row_entropy(logp_all) * weight).mean()
    policy_loss = -(logp * adv * weight).mean()
    value_loss = (torch.nn.functional.mse_loss(return_, value, reduction='none') * weight).mean()
    return policy_loss, value_loss, entropy_loss


def ppo_policy_error(logit_new, logit_old, action, adv, weight=None, logit_pretrained=None, clip_ratio: float = 0.2,
                     dual_clip: Optional[float] = None, entropy_bonus: bool = True, kl_type: str = 'k1'):
    """ppo.py:143-230 -> ``(policy_loss, entropy_loss, kl_div, approx_kl, clipfrac)``: the policy part of ``ppo_error``."""
    zero = torch.zeros_like(adv)
    p, _, e, k, approx_kl, clipfrac = ppo_error(logit_new, logit_old, action, zero, zero, adv, zero, weight,
                                                logit_pretrained, clip_ratio, False, dual_clip, kl_type)
    if not entropy_bonus:  # ppo.py:202-203
        e = torch.tensor(0.0)
    return p, e, k, approx_kl, clipfrac


def ppo_value_error(value_new, value_old, return_, weight=None, clip_ratio: float = 0.2, use_value_clip: bool = True):
    """ppo.py:263-275"""
    if weight is None:
        weight = torch.ones_like(value_old)
    if use_value_clip:
        value_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        v1 = (return_ - value_new).pow(2)
        v2 = (return_ - value_clip).pow(2)
        return 0.5 * (torch.max(v1, v2) * weight).mean()
    return 0.5 * ((return_ - value_new).pow(2) * weight).mean()


def nstep_return(reward, next_value, done, gamma: Union[float, list], nstep: int, value_gamma=None):
    assert reward.shape[0] == nstep  # td.py:257
    if isinstance(gamma, float):  # td.py:260-273
        disc = torch.ones(nstep)
        for i in range(1, nstep):
            disc[i] = gamma * disc[i - 1]
        acc = reward.mul(_trailing_ones(disc, reward)).sum(0)
        if value_gamma is None:
            return acc + (gamma ** nstep) * next_value * (1 - done)
        if not isinstance(value_gamma, torch.Tensor):  # np.isscalar branch td.py:269-270
            value_gamma = torch.full_like(next_value, value_gamma)
        return acc + _trailing_ones(value_gamma, next_value) * next_value * (1 - _trailing_ones(done, next_value))
    if isinstance(gamma, list):  # NGU per-sample gamma, td.py:275-282
        disc = torch.ones([nstep + 1, done.shape[0]])
        g = torch.stack(gamma, dim=0)
        for i in range(1, nstep + 1):
            disc[i] = g * disc[i - 1]
        disc = _trailing_ones(disc, reward)
        acc = reward.mul(disc[:nstep]).sum(0)
        return acc + disc[nstep] * next_value * (1 - done)
    raise TypeError("The type of gamma should be float or list")  # td.py:284


def value_transform(x, eps: float = 1e-2):
    """value_rescale.py:19  h(x) = sign(x)(sqrt(|x|+1)-1) + eps*x"""
    return torch.sign(x) * (torch.sqrt(torch.abs(x) + 1) - 1) + eps * x


def value_inv_transform(x, eps: float = 1e-2):
    """value_rescale.py:34"""
    return torch.sign(x) * (((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)


# --------------------------------------------------------------------------------------------------------------
# td.py:649-719 q_nstep_td_error ; td.py:810-867 ..._with_rescale
# --------------------------------------------------------------------------------------------------------------
def q_nstep_td_error(
        q,
        next_n_q,
        action,
        next_n_action,
        reward,
        done,
        weight=None,
        gamma: Union[float, list] = 0.99,
        nstep: int = 1,
        cum_reward: bool = False,
        value_gamma=None,
        criterion=None
):
    """Returns ``(loss, td_error_per_sample)``; default criterion is elementwise squared error (td.py:655)."""
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    if weight is None:
        weight = torch.ones_like(reward)  # td.py:692-693 -- (n, B) ones: the mean then runs over n*B rows
    if action.dim() == 1 or action.dim() < q.dim():  # td.py:695-699
        action = action.unsqueeze(-1)
    elif action.dim() > 1:  # td.py:700-705 (MARL with already-expanded action)
        reward = reward.unsqueeze(-1)
        weight = weight.unsqueeze(-1)
        done = done.unsqueeze(-1)
        if value_gamma is not None:
            value_gamma = value_gamma.unsqueeze(-1)
    q_sa = q.gather(-1, action).squeeze(-1)  # td.py:707
    tq = next_n_q.gather(-1, next_n_action.unsqueeze(-1)).squeeze(-1)  # td.py:709
    if cum_reward:  # td.py:711-715
        if value_gamma is None:
            tq = reward + (gamma ** nstep) * tq * (1 - done)
        else:
            tq = reward + value_gamma * tq * (1 - done)
    else:
        tq = nstep_return(reward, tq, done, gamma, nstep, value_gamma)  # td.py:717
    per_sample = criterion(q_sa, tq.detach())
    return (per_sample * weight).mean(), per_sample


def q_nstep_td_error_with_rescale(
        q, next_n_q, action, next_n_action, reward, done, weight=None, gamma=0.99, nstep: int = 1, value_gamma=None,
        criterion=None
):
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    assert action.dim() == 1, action.shape  # td.py:854
    if weight is None:
        weight = torch.ones_like(action)  # td.py:855-856 -- int64 ones
    rows = torch.arange(action.shape[0])
    q_sa = q[rows, action]
    tq = next_n_q[rows, next_n_action]
    tq = value_inv_transform(tq)  # td.py:862
    tq = nstep_return(reward, tq, done, gamma, nstep, value_gamma)  # td.py:863
    tq = value_transform(tq)  # td.py:864
    per_sample = criterion(q_sa, tq.detach())
    return (per_sample * weight).mean(), per_sample


def bdq_nstep_td_error(q, next_n_q, action, next_n_action, reward, done, weight=None, gamma=0.99, nstep: int = 1,
                       cum_reward: bool = False, value_gamma=None, criterion=None):
    """td.py:722-789: branching dueling Q; q (B, D, N), action (B, D); per-sample error = mean over the D branches (:788)."""
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    if weight is None:
        weight = torch.ones_like(reward)  # td.py:771-772
    reward = reward.unsqueeze(-1)  # td.py:773-776
    done = done.unsqueeze(-1)
    if value_gamma is not None:
        value_gamma = value_gamma.unsqueeze(-1)
    q_sa = q.gather(-1, action.unsqueeze(-1)).squeeze(-1)  # td.py:778
    tq = next_n_q.gather(-1, next_n_action.unsqueeze(-1)).squeeze(-1)
    if cum_reward:  # td.py:781-785
        if value_gamma is None:
            tq = reward + (gamma ** nstep) * tq * (1 - done)
        else:
            tq = reward + value_gamma * tq * (1 - done)
    else:
        tq = nstep_return(reward, tq, done, gamma, nstep, value_gamma)  # td.py:787
    per_sample = criterion(q_sa, tq.detach())
    per_sample = per_sample.mean(-1)  # td.py:788
    return (per_sample * weight).mean(), per_sample


def q_nstep_td_error_sequence(q, next_n_q, action, next_n_action, reward, done, weight=None, value_gamma=None,
                              gamma=0.99, nstep: int = 1, rescale: bool = False, priority_mix: float = 0.9):
    """The learner loop around the operator in the recurrent Q policies (ding/policy/r2d2.py:347-369; ngu.py:330-360):
    q (T, B, N), reward (T, nstep, B), done / weight / value_gamma (T, B).  Returns (loss, priority, td (T, B))."""
    fn = q_nstep_td_error_with_rescale if rescale else q_nstep_td_error
    losses, errs, raw = [], [], []
    for t in range(q.shape[0]):  # r2d2.py:347
        l, e = fn(q[t], next_n_q[t], action[t], next_n_action[t], reward[t], done[t],
                  None if weight is None else weight[t], gamma, nstep,
                  value_gamma=None if value_gamma is None else value_gamma[t])
        losses.append(l)
        errs.append(e.abs())
        raw.append(e)
    loss = sum(losses) / (len(losses) + 1e-8)  # r2d2.py:364
    prio = priority_mix * torch.max(torch.stack(errs), dim=0)[0] + (1 - priority_mix) * (
        torch.sum(torch.stack(errs), dim=0) / (len(errs) + 1e-8))  # r2d2.py:367-369
    return loss, prio.detach(), torch.stack(raw)


# --------------------------------------------------------------------------------------------------------------
# td.py:26-72 q_1step_td_error ; td.py:529-573 v_1step_td_error ; td.py:579-617 v_nstep_td_error
# --------------------------------------------------------------------------------------------------------------
def q_1step_td_error(q, next_q, act, next_act, reward, done, weight=None, gamma: float = 0.99, criterion=None):
    """Returns the loss only (td.py:63-72)."""
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    assert len(act.shape) == 1, act.shape        # td.py:64
    assert len(reward.shape) == 1, reward.shape  # td.py:65
    rows = torch.arange(act.shape[0])
    if weight is None:
        weight = torch.ones_like(reward)
    q_sa = q[rows, act]
    tq = next_q[rows, next_act]
    tq = gamma * (1 - done) * tq + reward  # td.py:71
    return (criterion(q_sa, tq.detach()) * weight).mean()


def v_1step_td_error(v, next_v, reward, done=None, weight=None, gamma: float = 0.99, criterion=None):
    """Returns ``(loss, td_error_per_sample)`` (td.py:557-573)."""
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    if weight is None:
        weight = torch.ones_like(v)
    if len(v.shape) == len(reward.shape):  # td.py:560-564
        target_v = gamma * (1 - done) * next_v + reward if done is not None else gamma * next_v + reward
    else:  # td.py:565-569
        if done is not None:
            target_v = gamma * (1 - done).unsqueeze(1) * next_v + reward.unsqueeze(1)
        else:
            target_v = gamma * next_v + reward.unsqueeze(1)
    per_sample = criterion(v, target_v.detach())
    return (per_sample * weight).mean(), per_sample


def v_nstep_td_error(v, next_n_v, reward, done, weight=None, value_gamma=None, gamma: float = 0.99, nstep: int = 1,
                     criterion=None):
    """Returns ``(loss, td_error_per_sample)`` (td.py:611-617)."""
    if criterion is None:
        criterion = torch.nn.MSELoss(reduction='none')
    if weight is None:
        weight = torch.ones_like(v)
    target_v = nstep_return(reward, next_n_v, done, gamma, nstep, value_gamma)
    per_sample = criterion(v, target_v.detach())
    return (per_sample * weight).mean(), per_sample


# --------------------------------------------------------------------------------------------------------------
# td.py:413-523 dist_nstep_td_error (C51 categorical projection)
# --------------------------------------------------------------------------------------------------------------
def dist_nstep_td_error(
        dist,
        next_n_dist,
        act,
        next_n_act,
        reward,
        done,
        weight=None,
        gamma: float = 0.99,
        v_min: float = -10.,
        v_max: float = 10.,
        n_atom: int = 51,
        nstep: int = 1,
        value_gamma=None
):
    """Returns ``(loss, td_error_per_sample)``; single-agent ``act (B,)`` and multi-agent ``act (B,A)``."""
    disc = torch.ones(nstep)
    for i in range(1, nstep):
        disc[i] = gamma * disc[i - 1]
    ret = torch.matmul(disc, reward)  # td.py:456
    support = torch.linspace(v_min, v_max, n_atom)  # td.py:457
    delta_z = (v_max - v_min) / (n_atom - 1)
    if act.dim() == 1:  # td.py:459-469
        ret = ret.unsqueeze(-1)
        done = done.unsqueeze(-1)
        nrow = act.shape[0]
        rows = torch.arange(nrow)
        if weight is None:
            weight = torch.ones_like(ret)
        elif isinstance(weight, float):
            weight = torch.tensor(weight)
        nd = next_n_dist[rows, next_n_act].detach()
    else:  # td.py:470-489
        n_b, n_a = act.shape
        ret = ret.unsqueeze(-1).repeat(1, n_a)
        done = done.unsqueeze(-1).repeat(1, n_a)
        nrow = n_b * n_a
        rows = torch.arange(nrow)
        n_act = dist.shape[2]
        dist = dist.reshape(nrow, n_act, -1)
        ret = ret.reshape(nrow, -1)
        done = done.reshape(nrow, -1)
        next_n_dist = next_n_dist.reshape(nrow, n_act, -1)
        next_n_act = next_n_act.reshape(nrow)
        nd = next_n_dist[rows, next_n_act].detach().reshape(nrow, -1)
        act = act.reshape(nrow)
        if weight is None:
            weight = torch.ones_like(ret)
        elif isinstance(weight, float):
            weight = torch.tensor(weight)
    if value_gamma is None:  # td.py:491-498
        tz = ret + (1 - done) * (gamma ** nstep) * support
    elif isinstance(value_gamma, float):
        tz = ret + (1 - done) * torch.tensor(value_gamma).unsqueeze(-1) * support
    else:
        tz = ret + (1 - done) * value_gamma.unsqueeze(-1) * support
    tz = tz.clamp(min=v_min, max=v_max)
    pos = (tz - v_min) / delta_z  # td.py:500
    lo = pos.floor().long()
    hi = pos.ceil().long()
    lo[(hi > 0) * (lo == hi)] -= 1  # td.py:504
    hi[(lo < (n_atom - 1)) * (lo == hi)] += 1  # td.py:505
    proj = torch.zeros_like(nd)
    base = torch.linspace(0, (nrow - 1) * n_atom, nrow).unsqueeze(1).expand(nrow, n_atom).long()  # td.py:508-509
    proj.view(-1).index_add_(0, (lo + base).view(-1), (nd * (hi.float() - pos)).view(-1))
    proj.view(-1).index_add_(0, (hi + base).view(-1), (nd * (pos - lo.float())).view(-1))
    picked = dist[rows, act]
    assert (picked > 0.0).all(), ("dist act", picked, "dist:", dist)  # td.py:513
    log_p = torch.log(picked)
    if weight.dim() == 1:
        weight = weight.unsqueeze(-1)
    per_sample = -(log_p * proj).sum(-1)  # td.py:519 (unweighted)
    loss = -(log_p * proj * weight).sum(-1).mean()  # td.py:521 (weighted)
    return loss, per_sample


def dist_1step_td_error(dist, next_dist, act, next_act, reward, done, weight=None, gamma: float = 0.99,
                        v_min: float = -10., v_max: float = 10., n_atom: int = 51):
    """td.py:294-383: C51 1-step; same projection as the n-step form with target_z = r + (1-done)*gamma*z; returns the loss
    only; no positivity assert."""
    assert len(reward.shape) == 1, reward.shape  # td.py:343
    support = torch.linspace(v_min, v_max, n_atom)
    delta_z = (v_max - v_min) / (n_atom - 1)
    if act.dim() == 1:  # td.py:347-354
        reward = reward.unsqueeze(-1)
        done = done.unsqueeze(-1)
        nrow = act.shape[0]
        rows = torch.arange(nrow)
        if weight is None:
            weight = torch.ones_like(reward)
        nd = next_dist[rows, next_act].detach()
    else:  # td.py:355-371
        n_b, n_a = act.shape
        reward = reward.unsqueeze(-1).repeat(1, n_a)
        done = done.unsqueeze(-1).repeat(1, n_a)
        nrow = n_b * n_a
        rows = torch.arange(nrow)
        n_act = dist.shape[2]
        dist = dist.reshape(nrow, n_act, -1)
        reward = reward.reshape(nrow, -1)
        done = done.reshape(nrow, -1)
        next_dist = next_dist.reshape(nrow, n_act, -1)
        next_act = next_act.reshape(nrow)
        nd = next_dist[rows, next_act].detach().reshape(nrow, -1)
        act = act.reshape(nrow)
        if weight is None:
            weight = torch.ones_like(reward)
    tz = reward + (1 - done) * gamma * support  # td.py:372
    tz = tz.clamp(min=v_min, max=v_max)
    pos = (tz - v_min) / delta_z
    lo = pos.floor().long()
    hi = pos.ceil().long()
    lo[(hi > 0) * (lo == hi)] -= 1
    hi[(lo < (n_atom - 1)) * (lo == hi)] += 1
    proj = torch.zeros_like(nd)
    offset = torch.linspace(0, (nrow - 1) * n_atom, nrow).unsqueeze(1).expand(nrow, n_atom).long()
    proj.view(-1).index_add_(0, (lo + offset).view(-1), (nd * (hi.float() - pos)).view(-1))
    proj.view(-1).index_add_(0, (hi + offset).view(-1), (nd * (pos - lo.float())).view(-1))
    log_p = torch.log(dist[rows, act])
    return -(log_p * proj * weight).sum(-1).mean()  # td.py:382


# --------------------------------------------------------------------------------------------------------------
# td.py:1539-1651 td_lambda_error / generalized_lambda_returns / multistep_forward_view
# --------------------------------------------------------------------------------------------------------------
def generalized_lambda_returns(bootstrap_values, rewards, gammas, lambda_, done=None):
    if not isinstance(gammas, torch.Tensor):
        gammas = gammas * torch.ones_like(rewards)
    if not isinstance(lambda_, torch.Tensor):
        lambda_ = lambda_ * torch.ones_like(rewards)
    nxt = bootstrap_values[1:, :]  # V_{t+1}, td.py:1604
    out = torch.empty_like(rewards)
    if done is None:
        done = torch.zeros_like(rewards)
    out[-1, :] = rewards[-1, :] + (1 - done[-1, :]) * gammas[-1, :] * nxt[-1, :]  # td.py:1642
    trace = gammas * lambda_
    for t in range(rewards.size(0) - 2, -1, -1):  # td.py:1644-1649
        out[t, :] = rewards[t, :] + (1 - done[t, :]) * (trace[t, :] * out[t + 1, :] + (gammas[t, :] - trace[t, :]) * nxt[t, :])
    return out


def lambda_returns_functional(bootstrap_values, rewards, gammas, lambda_, done=None):
    """The same recurrence written out of place (list + stack), so that autograd can differentiate it w.r.t. EVERY operand,
    including tensor gammas / lambdas -- the reference's in-place loop (td.py:1642-1649) supports gradients for
    bootstrap_values and rewards only.  Checker for the product's transposed-scan backward."""
    if not isinstance(gammas, torch.Tensor):
        gammas = gammas * torch.ones_like(rewards)
    if not isinstance(lambda_, torch.Tensor):
        lambda_ = lambda_ * torch.ones_like(rewards)
    if done is None:
        done = torch.zeros_like(rewards)
    nxt = bootstrap_values[1:]
    T = rewards.shape[0]
    rows = [None] * T
    rows[T - 1] = rewards[T - 1] + (1 - done[T - 1]) * gammas[T - 1] * nxt[T - 1]
    trace = gammas * lambda_
    for t in range(T - 2, -1, -1):
        rows[t] = rewards[t] + (1 - done[t]) * (trace[t] * rows[t + 1] + (gammas[t] - trace[t]) * nxt[t])
    return torch.stack(rows, 0)


def td_lambda_error(value, reward, weight=None, gamma: float = 0.9, lambda_: float = 0.8):
    if weight is None:
        weight = torch.ones_like(reward)
    with torch.no_grad():
        target = generalized_lambda_returns(value, reward, gamma, lambda_)
    err = torch.nn.functional.mse_loss(target, value[:-1], reduction='none')
    return 0.5 * (err * weight).mean()  # td.py:1570


# --------------------------------------------------------------------------------------------------------------
# upgo.py:7-111
# --------------------------------------------------------------------------------------------------------------
def upgo_returns(rewards, bootstrap_values):
    keep = (rewards + bootstrap_values[1:]) >= bootstrap_values[:-1]  # upgo.py:66
    keep = torch.cat([keep[1:], torch.ones_like(keep[-1:])], dim=0)  # upgo.py:67
    return generalized_lambda_returns(bootstrap_values, rewards, 1.0, keep)


def tb_cross_entropy(logit, label, mask=None):
    """upgo.py:7-43: NEGATIVE cross entropy (i.e. log-prob of the label), (T,B) out."""
    n_t, n_b = label.shape[:2]
    ce_fn = torch.nn.functional.cross_entropy
    if label.dim() > 2:
        assert label.dim() == 3
        s, n = logit.shape[-2:]
        ce = -ce_fn(logit.reshape(-1, n), label.reshape(-1), reduction='none').view(n_t * n_b, -1)
        if mask is not None:
            ce = ce * mask.reshape(-1, s)
        return ce.sum(dim=1).reshape(n_t, n_b)
    ce = -ce_fn(logit.reshape(-1, logit.shape[-1]), label.reshape(-1), reduction='none')
    return ce.reshape(n_t, n_b, -1).mean(dim=2)


def upgo_loss(target_output, rhos, action, rewards, bootstrap_values, mask=None):
    with torch.no_grad():
        g = upgo_returns(rewards, bootstrap_values)
        advantages = rhos * (g - bootstrap_values[:-1])
    metric = tb_cross_entropy(target_output, action, mask)
    assert metric.shape == action.shape[:2]
    return -(advantages * metric).mean()


# --------------------------------------------------------------------------------------------------------------
# vtrace.py:9-136 ; isw.py:55-58
# --------------------------------------------------------------------------------------------------------------
def vtrace_error_discrete_action(
        target_output,
        behaviour_output,
        action,
        value,
        reward,
        weight=None,
        gamma: float = 0.99,
        lambda_: float = 0.95,
        rho_clip_ratio: float = 1.0,
        c_clip_ratio: float = 1.0,
        rho_pg_clip_ratio: float = 1.0
):
    """Returns ``(policy_loss, value_loss, entropy_loss)``."""
    with torch.no_grad():
        lp_t = _chosen(_log_softmax_rows(target_output), action)
        lp_b = _chosen(_log_softmax_rows(behaviour_output), action)
        isw = torch.exp(lp_t - lp_b)  # isw.py:55-58
        rho = torch.clamp(isw, max=rho_clip_ratio)
        cs = torch.clamp(isw, max=c_clip_ratio)
        deltas = rho * (reward + gamma * value[1:] - value[:-1])  # vtrace.py:22
        trace = gamma * lambda_
        vs = value[:-1].clone()
        carry = 0.
        for t in range(reward.size(0) - 1, -1, -1):  # vtrace.py:26-28
            carry = deltas[t] + trace * cs[t] * carry
            vs[t] += carry
        rho_pg = torch.clamp(isw, max=rho_pg_clip_ratio)
        vs_next = torch.cat([vs[1:], value[-1:]], 0)  # vtrace.py:127
        adv = rho_pg * (reward + gamma * vs_next - value[:-1])  # vtrace.py:45
    if weight is None:
        weight = torch.ones_like(reward)
    lp_all = _log_softmax_rows(target_output)
    pg_loss = -(_chosen(lp_all, action) * adv * weight).mean()
    value_loss = (torch.nn.functional.mse_loss(value[:-1], vs, reduction='none') * weight).mean()  # no 0.5
    entropy_loss = (_row_entropy(lp_all) * weight).mean()
    return pg_loss, value_loss, entropy_loss


def vtrace_error_continuous_action(mu_target, sigma_target, mu_behaviour, sigma_behaviour, action, value, reward, weight=None,
                                   gamma: float = 0.99, lambda_: float = 0.95, rho_clip_ratio: float = 1.0,
                                   c_clip_ratio: float = 1.0, rho_pg_clip_ratio: float = 1.0):
    """vtrace.py:139-212 with the Independent(Normal) log-prob / entropy written out (see ppo_error_continuous)."""

    def logp(mu, sigma):
        return (-((action - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - math.log(math.sqrt(2 * math.pi))).sum(-1)

    with torch.no_grad():
        IS = torch.exp(logp(mu_target, sigma_target) - logp(mu_behaviour, sigma_behaviour))  # isw.py:49-53
        rhos = torch.clamp(IS, max=rho_clip_ratio)
        cs = torch.clamp(IS, max=c_clip_ratio)
        deltas = rhos * (reward + gamma * value[1:] - value[:-1])  # vtrace.py:22
        trace = gamma * lambda_
        return_ = value[:-1].clone()
        carry = 0.
        for t in range(reward.size(0) - 1, -1, -1):  # vtrace.py:26-28
            carry = deltas[t] + trace * cs[t] * carry
            return_[t] += carry
        pg_rhos = torch.clamp(IS, max=rho_pg_clip_ratio)
        return_t_plus_1 = torch.cat([return_[1:], value[-1:]], 0)
        adv = pg_rhos * (reward + gamma * return_t_plus_1 - value[:-1])  # vtrace.py:32-45
    if weight is None:
        weight = torch.ones_like(reward)
    pg_loss = -(logp(mu_target, sigma_target) * adv * weight).mean()
    value_loss = (torch.nn.functional.mse_loss(value[:-1], return_, reduction='none') * weight).mean()
    entropy_loss = ((0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma_target)).sum(-1) * weight).mean()
    return pg_loss, value_loss, entropy_loss


def _quantile_target(next_theta, reward, done, gamma, nstep, value_gamma):
    """n-step target of the three quantile heads (td.py:1144-1159): ``reward_factor`` built in fp32 by repeated multiplication,
    ``matmul`` with the (nstep, B) rewards, ``gamma ** nstep`` in python double."""
    reward_factor = torch.ones(nstep)
    for i in range(1, nstep):
        reward_factor[i] = gamma * reward_factor[i - 1]
    r = torch.matmul(reward_factor, reward)  # (B,)
    g = (gamma ** nstep) if value_gamma is None else value_gamma.reshape(-1, 1)
    return r.unsqueeze(-1) + g * next_theta * (1 - done).unsqueeze(-1)  # (B, n')


def _quantile_loss(theta, target, tau, weight, huber, strict, divisor, mean_over_target):
    """theta (B, n), target (B, n'), tau (B, n); u = target_j - theta_i, rho = |tau_i - 1[u <= 0 or < 0]| huber(u) / divisor.
    QR-DQN lays u out as (B, n, n') and takes sum(-1).mean(1) (td.py:1162-1164); IQN / FQF lay it out as (B, n', n, 1) and take
    sum(dim=2).mean(dim=1)[:, 0] (td.py:1325-1344, :1419-1434) -- kept, so that the summation order is the reference's."""
    if mean_over_target:
        u = target.unsqueeze(-1)[:, :, None, :] - theta.unsqueeze(-1)[:, None, :, :]  # (B, n', n, 1)
        ind = ((u < 0) if strict else (u <= 0)).float().detach()
        t = tau.unsqueeze(-1)[:, None, :, :].repeat([1, target.shape[1], 1, 1])
        loss = ((torch.abs(t - ind) * huber(u)) / divisor).sum(dim=2).mean(dim=1)[:, 0]
    else:
        u = target.unsqueeze(1) - theta.unsqueeze(2)  # (B, n, n')
        ind = ((u < 0) if strict else (u <= 0)).float().detach()
        loss = (huber(u) * (tau.unsqueeze(2) - ind).abs()).sum(-1).mean(1)
    if weight is None:
        weight = torch.ones_like(loss)
    return (loss * weight).mean(), loss


def _smooth_l1(u):
    return torch.where(u.abs() < 1.0, 0.5 * u * u, u.abs() - 0.5)


def qrdqn_nstep_td_error(q, next_n_q, action, next_n_action, reward, done, tau, weight=None, gamma: float = 0.99,
                         nstep: int = 1, value_gamma=None):
    """td.py:1098-1166: q / next_n_q (B, N, num), smooth-l1, indicator u <= 0, sum over the target axis, mean over num."""
    rows = torch.arange(action.shape[0])
    theta = q[rows, action]                # (B, num)
    target = _quantile_target(next_n_q[rows, next_n_action], reward, done, gamma, nstep, value_gamma)
    B, num = theta.shape
    t = torch.broadcast_to(torch.as_tensor(tau, dtype=torch.float32), (B, num, target.shape[1]))[:, :, 0]
    return _quantile_loss(theta, target, t, weight, _smooth_l1, False, 1.0, False)


def iqn_nstep_td_error(q, next_n_q, action, next_n_action, reward, done, replay_quantiles, weight=None, gamma: float = 0.99,
                       nstep: int = 1, kappa: float = 1.0, value_gamma=None):
    """td.py:1253-1346: q (tau, B, N), next_n_q (tau', B, N), Huber(kappa) through torch.where(|u| <= kappa), indicator u < 0,
    / kappa, sum over tau, mean over tau'."""
    tau_n, B = q.shape[0], done.shape[0]
    rows = torch.arange(B)
    theta = q[:, rows, action].t()         # (B, tau)
    target = _quantile_target(next_n_q[:, rows, next_n_action].t(), reward, done, gamma, nstep, value_gamma)

    def huber(u):
        return torch.where(u.abs() <= kappa, 0.5 * u ** 2, kappa * (u.abs() - 0.5 * kappa))

    return _quantile_loss(theta, target, replay_quantiles.reshape(tau_n, B).t(), weight, huber, True, kappa, True)


def fqf_nstep_td_error(q, next_n_q, action, next_n_action, reward, done, quantiles_hats, weight=None, gamma: float = 0.99,
                       nstep: int = 1, kappa: float = 1.0, value_gamma=None):
    """td.py:1359-1436: q (B, tau, N), next_n_q (B, tau', N), smooth-l1 (beta 1), indicator u < 0, / kappa, sum over tau, mean
    over tau'."""
    rows = torch.arange(action.shape[0])
    theta = q[rows, :, action]             # (B, tau)
    target = _quantile_target(next_n_q[rows, :, next_n_action], reward, done, gamma, nstep, value_gamma)
    return _quantile_loss(theta, target, quantiles_hats, weight, _smooth_l1, True, kappa, True)


def compute_q_retraces(q_values, v_pred, rewards, actions, weights, ratio, gamma: float = 0.9):
    """retrace.py:7-56: Qret[T] = V[T]; Qret[t] = r_t + gamma w_t tmp; tmp = min(ratio_t[a_t], 1) (Qret[t] - Q_t[a_t]) + V_t."""
    T = q_values.size()[0] - 1
    rewards, actions, weights = rewards.unsqueeze(-1), actions.unsqueeze(-1), weights.unsqueeze(-1)
    q_retraces = torch.zeros_like(v_pred)
    tmp = v_pred[-1]
    q_retraces[-1] = v_pred[-1]
    q_gather = q_values[0:-1].gather(-1, actions)
    ratio_gather = ratio.gather(-1, actions)
    for idx in reversed(range(T)):
        q_retraces[idx] = rewards[idx] + gamma * weights[idx] * tmp
        tmp = ratio_gather[idx].clamp(max=1.0) * (q_retraces[idx] - q_gather[idx]) + v_pred[idx]
    return q_retraces


def happo_error(logit_new, logit_old, action, value_new, value_old, adv, return_, weight=None, factor=None,
                clip_ratio: float = 0.2, use_value_clip: bool = True, dual_clip: Optional[float] = None):
    """happo.py:18-78 (= happo_policy_error :81-147 + happo_value_error :150-192): ppo_error whose selected surrogate is
    multiplied by ``factor.squeeze(1)`` before the dual clip.  Returns (policy_loss, value_loss, entropy_loss, approx_kl,
    clipfrac)."""
    assert dual_clip is None or dual_clip > 1.0
    w = torch.ones_like(adv) if weight is None else weight
    lp_new_all = _log_softmax_rows(logit_new)
    lp_new = _chosen(lp_new_all, action)
    lp_old = _chosen(_log_softmax_rows(logit_old), action)
    ent = _row_entropy(lp_new_all)
    if ent.shape != w.shape:  # happo.py:114-115
        ent = ent.mean(dim=1)
    entropy_loss = (ent * w).mean()
    ratio = torch.exp(lp_new - lp_old)
    if ratio.shape != adv.shape:
        ratio = ratio.mean(dim=1)
    surr1 = ratio * adv
    surr2 = ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv
    clip1 = torch.min(surr1, surr2) * factor.squeeze(1)  # happo.py:125
    if dual_clip is not None:
        clip2 = torch.max(clip1, dual_clip * adv)
        policy_loss = -(torch.where(adv < 0, clip2, clip1) * w).mean()
    else:
        policy_loss = (-clip1 * w).mean()
    with torch.no_grad():
        approx_kl = (lp_old - lp_new).mean().item()
        clipfrac = (ratio.gt(1 + clip_ratio) | ratio.lt(1 - clip_ratio)).float().mean().item()
    wv = torch.ones_like(value_old) if weight is None else weight
    if use_value_clip:
        value_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        value_loss = 0.5 * (torch.max((return_ - value_new).pow(2), (return_ - value_clip).pow(2)) * wv).mean()
    else:
        value_loss = 0.5 * ((return_ - value_new).pow(2) * wv).mean()
    return policy_loss, value_loss, entropy_loss, approx_kl, clipfrac


def acer_policy_error(q_values, q_retraces, v_pred, target_logit, actions, ratio, c_clip_ratio: float = 10.0):
    """acer.py:8-57 -> (actor_loss, bias_correction_loss), both (T, B, 1)."""
    actions = actions.unsqueeze(-1)
    with torch.no_grad():
        advantage_retraces = q_retraces - v_pred
        advantage_native = q_values - v_pred
    actor_loss = ratio.gather(-1, actions).clamp(max=c_clip_ratio) * advantage_retraces * target_logit.gather(-1, actions)
    bias = (1.0 - c_clip_ratio / (ratio + 1e-8)).clamp(min=0.0) * torch.exp(target_logit).detach() * advantage_native * \
        target_logit
    return actor_loss, bias.sum(-1, keepdim=True)


def acer_value_error(q_values, q_retraces, actions):
    """acer.py:60-83."""
    return 0.5 * (q_retraces - q_values.gather(-1, actions.unsqueeze(-1))).pow(2)


def acer_trust_region_update(actor_gradients, target_logit, avg_logit, trust_region_value):
    """acer.py:86-124."""
    with torch.no_grad():
        k = torch.exp(avg_logit)
    g = actor_gradients[0]
    scale = g.mul(k).sum(-1, keepdim=True) - trust_region_value
    scale = torch.div(scale, k.mul(k).sum(-1, keepdim=True)).clamp(min=0.0)
    return [g - scale * k]


def happo_error_continuous(mu_new, sigma_new, mu_old, sigma_old, action, value_new, value_old, adv, return_, weight=None,
                           factor=None, clip_ratio: float = 0.2, use_value_clip: bool = True,
                           dual_clip: Optional[float] = None):
    """happo.py:195-284 with Normal.log_prob / entropy written out (per dimension, NOT summed: the reference uses Normal, not
    Independent(Normal), so the entropy mean and approx_kl run over the B x D terms).  Returns (policy_loss, value_loss,
    entropy_loss, approx_kl, clipfrac)."""
    assert dual_clip is None or dual_clip > 1.0
    if weight is None:
        weight = torch.ones_like(adv)

    def logp(mu, sigma):
        return -((action - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - math.log(math.sqrt(2 * math.pi))

    if mu_old.dim() == 1:  # happo.py:237-238
        mu_old, sigma_old = mu_old.unsqueeze(-1), sigma_old.unsqueeze(-1)
    logp_new, logp_old = logp(mu_new, sigma_new), logp(mu_old, sigma_old)
    entropy_loss = ((0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma_new)) * weight.unsqueeze(1)).mean()
    ratio = torch.prod(torch.exp(logp_new - logp_old), dim=-1)
    surr1 = ratio * adv
    surr2 = ratio.clamp(1 - clip_ratio, 1 + clip_ratio) * adv
    if dual_clip is not None:
        policy_loss = (-torch.max(factor.squeeze(1) * torch.min(surr1, surr2), dual_clip * adv) * weight).mean()
    else:
        policy_loss = (-factor.squeeze(1) * torch.min(surr1, surr2) * weight).mean()
    with torch.no_grad():
        approx_kl = (logp_old - logp_new).mean().item()
        clipfrac = (ratio.gt(1 + clip_ratio) | ratio.lt(1 - clip_ratio)).float().mean().item()
    if use_value_clip:
        value_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        value_loss = 0.5 * (torch.max((return_ - value_new).pow(2), (return_ - value_clip).pow(2)) * weight).mean()
    else:
        value_loss = 0.5 * ((return_ - value_new).pow(2) * weight).mean()
    return policy_loss, value_loss, entropy_loss, approx_kl, clipfrac


def ppg_joint_error(logit_new, logit_old, action, value_new, value_old, return_, weight=None, clip_ratio: float = 0.2,
                    use_value_clip: bool = True):
    """ppg.py:10-69 -> (auxiliary_loss, behavioral_cloning_loss).  The second is F.kl_div(logp_new, logp_old, 'batchmean') with a
    log-probability where kl_div expects a probability: NaN value, finite gradient -- restated as written."""
    if weight is None:
        weight = torch.ones_like(return_)
    if use_value_clip:
        value_clip = value_old + (value_new - value_old).clamp(-clip_ratio, clip_ratio)
        auxiliary_loss = 0.5 * (torch.max((return_ - value_new).pow(2), (return_ - value_clip).pow(2)) * weight).mean()
    else:
        auxiliary_loss = 0.5 * ((return_ - value_new).pow(2) * weight).mean()
    logp_new = _chosen(_log_softmax_rows(logit_new), action)
    logp_old = _chosen(_log_softmax_rows(logit_old), action)
    return auxiliary_loss, torch.nn.functional.kl_div(logp_new, logp_old, reduction='batchmean')


# PARITY UNPINNED for the two functions below: ding.policy.common_utils / ding.utils.data import treetensor, easydict, gym ...,
# none of which is installed here, so the live functions cannot be run next to this restatement (unlike everything above).
def default_collate_flat(batch, cat_1dim=True):
    """ding/utils/data/collate_fn.py:80-160 for the field types a transition dict holds: tensors ((1,) samples are concatenated
    when cat_1dim, :133-135), numpy arrays, python floats (-> float32), ints (-> int64), bools, nested dicts."""
    elem = batch[0]
    if isinstance(elem, torch.Tensor):
        if elem.shape == (1, ) and cat_1dim:
            return torch.cat(batch, 0)
        return torch.stack(batch, 0)
    if type(elem).__module__ == 'numpy':
        if type(elem).__name__ == 'ndarray':
            return default_collate_flat([torch.as_tensor(b) for b in batch], cat_1dim)
        return torch.as_tensor(batch)
    if isinstance(elem, bool):
        return torch.tensor(batch)
    if isinstance(elem, float):
        return torch.tensor(batch, dtype=torch.float32)
    if isinstance(elem, int):
        return torch.tensor(batch, dtype=torch.int64)
    if isinstance(elem, dict):
        return {k: default_collate_flat([d[k] for d in batch], cat_1dim) for k in elem if not str(k).startswith('collate_ignore')}
    if elem is None:
        return None
    raise TypeError(type(elem))


def default_preprocess_learn(data, use_priority_IS_weight=False, use_priority=False, use_nstep=False, ignore_done=False):
    """ding/policy/common_utils.py:28-98, line for line on top of the collate restatement above."""
    elem = data[0]
    if isinstance(elem['action'], (np.ndarray, torch.Tensor)) and elem['action'].dtype in [np.int64, torch.int64]:
        data = default_collate_flat(data, cat_1dim=True)
    else:
        data = default_collate_flat(data, cat_1dim=False)
    if 'value' in data and data['value'].dim() == 2 and data['value'].shape[1] == 1:
        data['value'] = data['value'].squeeze(-1)
    if 'adv' in data and data['adv'].dim() == 2 and data['adv'].shape[1] == 1:
        data['adv'] = data['adv'].squeeze(-1)
    data['done'] = torch.zeros_like(data['done']).float() if ignore_done else data['done'].float()
    if data['done'].dim() == 2 and data['done'].shape[1] == 1:
        data['done'] = data['done'].squeeze(-1)
    if use_priority_IS_weight:
        assert use_priority, "Use IS Weight correction, but Priority is not used."
    if use_priority and use_priority_IS_weight:
        data['weight'] = data['priority_IS'] if 'priority_IS' in data else data['IS']
    else:
        data['weight'] = data.get('weight', None)
    if use_nstep:
        reward = data['reward']
        if len(reward.shape) == 1:
            reward = reward.unsqueeze(1)
        if reward.ndim == 2:
            data['reward'] = reward.transpose(0, 1).contiguous()
        elif reward.ndim == 3:
            data['reward'] = reward.permute(2, 0, 1).contiguous()
        else:
            raise ValueError("The 'reward' tensor must be either 2D or 3D. Got shape: {}".format(reward.shape))
    else:
        if data['reward'].dim() == 2 and data['reward'].shape[1] == 1:
            data['reward'] = data['reward'].squeeze(-1)
    return data


def adder_get_nstep_return_data(data, nstep, cum_reward=False, correct_terminate_gamma=True, gamma=0.99):
    """ding/rl_utils/adder.py:97-155, line for line on a list of per-step dicts (PARITY UNPINNED: adder.py imports ding.utils)."""
    if nstep == 1:
        return data
    fake_reward = torch.zeros_like(data[0]['reward'])
    next_obs_flag = 'next_obs' in data[0]
    for i in range(len(data) - nstep):
        if next_obs_flag:
            data[i]['next_obs'] = data[i + nstep]['obs']
        if cum_reward:
            data[i]['reward'] = sum([data[i + j]['reward'] * (gamma ** j) for j in range(nstep)])
        else:
            data[i]['reward'] = torch.cat([data[i + j]['reward'] for j in range(nstep)], dim=-1)
        data[i]['done'] = data[i + nstep - 1]['done']
        if correct_terminate_gamma:
            data[i]['value_gamma'] = gamma ** nstep
    for i in range(max(0, len(data) - nstep), len(data)):
        if next_obs_flag:
            data[i]['next_obs'] = data[-1]['next_obs']
        if cum_reward:
            data[i]['reward'] = sum([data[i + j]['reward'] * (gamma ** j) for j in range(len(data) - i)])
        else:
            data[i]['reward'] = torch.cat(
                [data[i + j]['reward'] for j in range(len(data) - i)] + [fake_reward for _ in range(nstep - (len(data) - i))], dim=-1
            )
        data[i]['done'] = data[-1]['done']
        if correct_terminate_gamma:
            data[i]['value_gamma'] = gamma ** (len(data) - i - 1)
    return data
