"""ACER heads with the signatures of ding/rl_utils/acer.py:8,60,86 (csrc/acer.cu): un-reduced per-transition losses, as
``ACERPolicy._forward_learn`` consumes them (policy/acer.py:247-270)."""
from typing import List, Tuple

import torch

from .. import ops

EPS = 1e-8




def acer_policy_error(
        q_values: torch.Tensor,
        q_retraces: torch.Tensor,
        v_pred: torch.Tensor,
        target_logit: torch.Tensor,
        actions: torch.Tensor,
        ratio: torch.Tensor,
        c_clip_ratio: float = 10.0
) -> Tuple[torch.Tensor, torch.Tensor]:
    """
    Drop-in for ding/rl_utils/acer.py:8-57.  q_values, target_logit (log pi), ratio (T, B, N); q_retraces, v_pred (T, B, 1);
    actions (T, B).  Returns ``(actor_loss, bias_correction_loss)``, both (T, B, 1) and attached to ``target_logit``
    (the advantages are formed under ``no_grad`` and ``exp(target_logit)`` is detached in the reference, acer.py:44-52; ``ratio``
    and the critic outputs are treated as data -- ACERPolicy computes them under ``no_grad`` / detaches them).
    """
    lead = tuple(actions.shape)
    N = target_logit.shape[-1]
    if tuple(target_logit.shape) != lead + (N, ) or tuple(q_values.shape) != lead + (N, ) or tuple(ratio.shape) != lead + (N, ) \
            or q_retraces.numel() != actions.numel() or v_pred.numel() != actions.numel():
        raise ValueError("acer_policy_error: q_values %s / q_retraces %s / v_pred %s / target_logit %s / actions %s / ratio %s" %
                         tuple(tuple(x.shape) for x in (q_values, q_retraces, v_pred, target_logit, actions, ratio)))
    dev = ops.compute_device(target_logit, q_values)
    host_out = not target_logit.is_cuda
    M = actions.numel()
    actor, bias = ops.AcerPolicyFunction.apply(
        ops.f32c(ops.to_device(target_logit, dev), 'target_logit'), ops.f32c(ops.to_device(q_values.detach(), dev), 'q_values'),
        ops.f32c(ops.to_device(q_retraces.detach(), dev), 'q_retraces'), ops.f32c(ops.to_device(v_pred.detach(), dev), 'v_pred'),
        ops.i64c(ops.to_device(actions, dev), N, 'actions'), ops.f32c(ops.to_device(ratio.detach(), dev), 'ratio'), M, N, float(c_clip_ratio)
    )
    actor, bias = actor.view(lead + (1, )), bias.view(lead + (1, ))
    return (actor.cpu(), bias.cpu()) if host_out else (actor, bias)


def acer_value_error(q_values, q_retraces, actions):
    """Drop-in for ding/rl_utils/acer.py:60-83: critic_loss (T, B, 1) = 0.5 (q_retraces - q_values[a])^2, attached to ``q_values``
    (``q_retraces`` is a no-grad target in ACERPolicy, policy/acer.py:231-232)."""
    lead = tuple(actions.shape)
    N = q_values.shape[-1]
    if tuple(q_values.shape) != lead + (N, ) or q_retraces.numel() != actions.numel():
        raise ValueError("acer_value_error: q_values %s / q_retraces %s / actions %s" %
                         (tuple(q_values.shape), tuple(q_retraces.shape), tuple(actions.shape)))
    dev = ops.compute_device(q_values)
    host_out = not q_values.is_cuda
    loss = ops.AcerValueFunction.apply(
        ops.f32c(ops.to_device(q_values, dev), 'q_values'), ops.f32c(ops.to_device(q_retraces.detach(), dev), 'q_retraces'),
        ops.i64c(ops.to_device(actions, dev), N, 'actions'), actions.numel(), N
    ).view(lead + (1, ))
    return loss.cpu() if host_out else loss


def acer_trust_region_update(
        actor_gradients: List[torch.Tensor], target_logit: torch.Tensor, avg_logit: torch.Tensor,
        trust_region_value: float
) -> List[torch.Tensor]:
    """Drop-in for ding/rl_utils/acer.py:86-124 (one gradient tensor in the list, as the reference; ``target_logit`` is unused
    there too)."""
    g = actor_gradients[0]
    dev = ops.compute_device(g, avg_logit)
    host_out = not g.is_cuda
    out = ops.acer_trust_region_(ops.f32c(ops.to_device(g.detach(), dev), 'actor_gradient'),
                                 ops.f32c(ops.to_device(avg_logit.detach(), dev), 'avg_logit'), trust_region_value)
    return [out.cpu() if host_out else out]
