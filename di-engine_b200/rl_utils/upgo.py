"""``upgo_loss`` / ``upgo_returns`` with the signatures of ding/rl_utils/upgo.py:46,77 -- csrc/td.cu + csrc/pg.cu."""
import torch

from .. import ops


def upgo_returns(rewards: torch.Tensor, bootstrap_values: torch.Tensor) -> torch.Tensor:
    """
    UPGO return targets (ding/rl_utils/upgo.py:46-68): a lambda-return with gamma = 1 whose trace continues
    (lambda_t = 1) while r_{t+1} + V_{t+2} >= V_{t+1}.  rewards (T, B), bootstrap_values (T+1, B) -> (T, B).
    The comparison and the recurrence are evaluated in one kernel, bit-exact with the reference.
    """
    dev = ops.compute_device(rewards, bootstrap_values)
    host_out = not rewards.is_cuda
    v = ops.f32c(ops.to_device(bootstrap_values.detach(), dev), 'bootstrap_values')
    r = ops.f32c(ops.to_device(rewards.detach(), dev), 'rewards')
    ret = ops.lambda_returns_(v, r, None, 1.0, None, 1.0, None, True)
    return ret.cpu() if host_out else ret


def upgo_loss(
        target_output: torch.Tensor,
        rhos: torch.Tensor,
        action: torch.Tensor,
        rewards: torch.Tensor,
        bootstrap_values: torch.Tensor,
        mask=None
) -> torch.Tensor:
    """
    Importance-weighted UPGO policy-gradient loss, drop-in for ding/rl_utils/upgo.py:77-111.
    target_output (T, B, N) with action (T, B) -- or (T, B, N2, N) with action / mask (T, B, N2) (upgo.py:25-37);
    rhos, rewards (T, B); bootstrap_values (T+1, B).  Gradient reaches ``target_output`` only.
    """
    dev = ops.compute_device(target_output, rewards)
    host_out = not target_output.is_cuda
    assert action.dim() >= 2  # upgo.py:23
    T, B = action.shape[:2]
    N = target_output.shape[-1]
    if action.dim() > 2:
        assert action.dim() == 3  # upgo.py:27
        K = action.shape[2]
    else:
        K = 1
    if target_output.numel() != T * B * K * N:
        raise ValueError("target_output %s does not match action %s" % (tuple(target_output.shape),
                                                                        tuple(action.shape)))
    logit = ops.f32c(ops.to_device(target_output, dev), 'target_output')
    act = ops.i64c(ops.to_device(action, dev))
    m = None
    if mask is not None and K > 1:  # the reference ignores mask for 3-D logits (upgo.py:38-42)
        m = ops.f32c(ops.to_device(mask.detach(), dev), 'mask')
    rho = ops.f32c(ops.to_device(rhos.detach(), dev), 'rhos')
    r = ops.f32c(ops.to_device(rewards.detach(), dev), 'rewards')
    v = ops.f32c(ops.to_device(bootstrap_values.detach(), dev), 'bootstrap_values')
    ret = ops.lambda_returns_(v, r, None, 1.0, None, 1.0, None, True)
    loss = ops.UPGOFunction.apply(logit, act, m, rho, ret, v, T * B, K, N)
    return loss.cpu() if host_out else loss
