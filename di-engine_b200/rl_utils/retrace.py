"""``compute_q_retraces`` with the signature of ding/rl_utils/retrace.py:7 -- one reverse-scan launch (csrc/retrace.cu)."""
import torch

from .. import _lib, ops


def compute_q_retraces(
        q_values: torch.Tensor,
        v_pred: torch.Tensor,
        rewards: torch.Tensor,
        actions: torch.Tensor,
        weights: torch.Tensor,
        ratio: torch.Tensor,
        gamma: float = 0.9
) -> torch.Tensor:
    """
    Drop-in for ding/rl_utils/retrace.py:7-56 (ACER, policy/acer.py:231-232).  q_values (T+1, B, N); v_pred (T+1, B, 1);
    rewards, actions, weights (T, B); ratio (T, B, N).  Returns q_retraces (T+1, B, 1), bit-identical to the reference loop.
    As in the reference the result carries no gradient.
    """
    T = q_values.size()[0] - 1
    if v_pred.dim() != 3 or v_pred.shape[-1] != 1 or v_pred.shape[0] != T + 1 or rewards.shape[0] != T \
            or tuple(rewards.shape) != tuple(actions.shape) or tuple(rewards.shape) != tuple(weights.shape) \
            or tuple(ratio.shape) != (T, ) + tuple(q_values.shape[1:]) or tuple(q_values.shape[:2]) != tuple(v_pred.shape[:2]):
        raise ValueError("compute_q_retraces: q_values %s / v_pred %s / rewards %s / actions %s / weights %s / ratio %s" % tuple(
            tuple(x.shape) for x in (q_values, v_pred, rewards, actions, weights, ratio)))
    dev = ops.compute_device(q_values, v_pred)
    host_out = not q_values.is_cuda
    B, N = q_values.shape[1], q_values.shape[2]
    q = ops.f32c(ops.to_device(q_values.detach(), dev), 'q_values')
    v = ops.f32c(ops.to_device(v_pred.detach(), dev), 'v_pred')
    r = ops.f32c(ops.to_device(rewards.detach(), dev), 'rewards')
    a = ops.i64c(ops.to_device(actions, dev), N, 'actions')
    w = ops.f32c(ops.to_device(weights.detach(), dev), 'weights')
    c = ops.f32c(ops.to_device(ratio.detach(), dev), 'ratio')
    out = torch.empty_like(v)
    with ops.on_device(dev):
        rc = ops.lib().b200rl_q_retraces(ops.ptr(q), ops.ptr(v), ops.ptr(r), ops.ptr(a), ops.ptr(w), ops.ptr(c), T, B, N,
                                         float(gamma), ops.ptr(out), ops.stream_ptr())
    _lib.check(rc, 'b200rl_q_retraces')
    return out.cpu() if host_out else out
