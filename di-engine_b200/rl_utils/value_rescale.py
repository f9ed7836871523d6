"""Value rescaling helpers h / h^-1 (ding/rl_utils/value_rescale.py:4-34).

Inside ``q_nstep_td_error_with_rescale`` the default pair is evaluated in the fused CUDA kernel (csrc/td.cu
``value_h`` / ``value_h_inv``); these tensor-level versions exist because they are the default ``trans_fn`` /
``inv_trans_fn`` arguments of that signature and part of the public namespace.
"""
import torch


def value_transform(x: torch.Tensor, eps: float = 1e-2) -> torch.Tensor:
    """h(x) = sign(x)(sqrt(|x|+1) - 1) + eps*x  (arXiv:1805.11593)."""
    root = torch.sqrt(x.abs() + 1) - 1
    return x.sign() * root + eps * x


def value_inv_transform(x: torch.Tensor, eps: float = 1e-2) -> torch.Tensor:
    """h^-1(x) = sign(x)(((sqrt(1 + 4 eps (|x| + 1 + eps)) - 1) / (2 eps))^2 - 1)."""
    inner = torch.sqrt(1 + 4 * eps * (x.abs() + 1 + eps))
    return x.sign() * (((inner - 1) / (2 * eps)) ** 2 - 1)
