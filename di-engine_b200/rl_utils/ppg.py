"""``ppg_joint_error`` with the signature of ding/rl_utils/ppg.py:10 -- the auxiliary value term on the PPO value kernel, the
behavioural-cloning term on its own small kernel (csrc/heads.cu)."""
from collections import namedtuple
from typing import Tuple

import torch

from .. import ops
from .ppo import ppo_value_data, ppo_value_error

ppg_data = namedtuple('ppg_data', ['logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'return_', 'weight'])
ppg_joint_loss = namedtuple('ppg_joint_loss', ['auxiliary_loss', 'behavioral_cloning_loss'])


def ppg_joint_error(
        data: namedtuple,
        clip_ratio: float = 0.2,
        use_value_clip: bool = True,
) -> Tuple[namedtuple, namedtuple]:
    """
    Drop-in for ding/rl_utils/ppg.py:10-69.  logit_new / logit_old (B, N), action (B,), value_new / value_old / return_ (B,),
    weight None or (B,).  ``auxiliary_loss`` is the clipped value loss of ``ppo_value_error`` (same expression, ppg.py:49-55).
    ``behavioral_cloning_loss`` is ``F.kl_div(logp_new, logp_old, reduction='batchmean')`` as the reference writes it -- the old
    LOG-probability in the place of a probability target: its value is NaN whenever an old log-probability is negative, its
    gradient ``-logp_old / B`` w.r.t. ``logp_new`` is finite; both are reproduced.
    """
    logit_new, logit_old, action, value_new, value_old, return_, weight = data
    aux = ppo_value_error(ppo_value_data(value_new, value_old, return_, weight), clip_ratio, use_value_clip)
    if logit_new.dim() != 2 or logit_old.shape != logit_new.shape or action.shape != logit_new.shape[:1]:
        raise ValueError("ppg_joint_error: logit_new %s / logit_old %s / action %s" %
                         (tuple(logit_new.shape), tuple(logit_old.shape), tuple(action.shape)))
    dev = ops.compute_device(logit_new, value_new)
    host_out = not logit_new.is_cuda
    bc = ops.ppg_bc_(ops.f32c(ops.to_device(logit_new, dev), 'logit_new'),
                     ops.f32c(ops.to_device(logit_old.detach(), dev), 'logit_old'), ops.i64c(ops.to_device(action, dev), logit_new.shape[-1]))
    return ppg_joint_loss(aux, bc.cpu() if host_out else bc)
