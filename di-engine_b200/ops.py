"""Tensor-level operators over the C ABI (``include/b200rl.h``): allocation of outputs, stream / workspace plumbing and
``torch.autograd.Function`` wrappers.  torch is used for device memory, streams and autograd bookkeeping only; every
arithmetic step of the hot path runs in the CUDA kernels of ``csrc/``.

Nothing here computes on the CPU.  Host (CPU) tensors are accepted by the public API in ``rl_utils`` by staging them
to the current CUDA device and returning results on the host -- the "host buffers" end-to-end path.
"""
import os

import torch

from . import _lib

_WS = {}
_CONST = {}


def lib():
    return _lib.load()


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.B200RLError(
            "di_engine_b200 needs a CUDA device (sm_100a). There is no CPU implementation of the operators; "
            "the CPU oracle under oracle/ is test infrastructure only."
        )
    lib()


def stream_ptr():
    """raw cudaStream_t of torch's current stream on the current device (the fast private accessor when torch has it: the
    public ``torch.cuda.current_stream().cuda_stream`` costs ~1.5 us of host time per call, and an operator asks twice)"""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


class _NoCtx:

    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NOCTX = _NoCtx()


def on_device(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (entering the guard costs several
    microseconds of host time -- more than the launch it protects at the learner's real batch sizes)"""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NOCTX
    return torch.cuda.device(device)


def workspace(device):
    """One zero-initialised scratch buffer per (device, stream): launches sharing it are stream-ordered."""
    idx = device.index
    if idx is None:
        idx = torch.cuda.current_device() if device.type == 'cuda' else -1
    key = (idx, stream_ptr())
    ws = _WS.get(key)
    if ws is None:
        nbytes = lib().b200rl_workspace_bytes()
        ws = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
        _WS[key] = ws
    return ws


def ptr(t):
    return None if t is None else t.data_ptr()


def f32c(t, name='tensor'):
    """fp32, contiguous; integer / bool flags (done, traj_flag, masks) are widened like the reference's ``.float()``."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        if t.dtype in (torch.float64, torch.float16, torch.bfloat16):
            raise TypeError("di_engine_b200: %s must be float32 (got %s); the B200 path computes in fp32" %
                            (name, t.dtype))
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# B200RL_CHECK_INDICES=1 (or ops.CHECK_INDICES = True): every action / label tensor is range-checked on the host before its
# pointer goes to a kernel -- an out-of-range index raises IndexError as torch's gather would, at the price of one device
# synchronisation per call.  Off by default: the kernels index with the value they are given.
CHECK_INDICES = os.environ.get('B200RL_CHECK_INDICES', '0') == '1'


def i64c(t, n=None, name='action'):
    if t.dtype != torch.int64:
        t = t.long()
    if CHECK_INDICES and n is not None and t.numel():
        lo, hi = torch.aminmax(t)
        lo, hi = int(lo), int(hi)
        if lo < 0 or hi >= n:
            raise IndexError("di_engine_b200: %s holds index %d, out of range for %d classes" % (name, lo if lo < 0 else hi, n))
    return t if t.is_contiguous() else t.contiguous()


def to_device(t, device):
    if isinstance(t, torch.Tensor) and t.device != device:
        return t.to(device, non_blocking=True)
    return t


def compute_device(*tensors):
    """Device the op runs on: the device of the first CUDA tensor, else the current CUDA device (host-buffer path)."""
    require_cuda()
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device('cuda', torch.cuda.current_device())


def const_scalar(value, device):
    """A cached 1-element device tensor holding a python scalar (weights / value_gamma given as floats)."""
    key = (float(value), device.index)
    t = _CONST.get(key)
    if t is None:
        t = torch.full((1, ), float(value), dtype=torch.float32, device=device)
        _CONST[key] = t
    return t


def _g(grad):
    """Upstream gradient of a 0-dim loss as a device pointer (None -> the kernel treats it as 0)."""
    if grad is None:
        return None, None
    g = grad if grad.dtype == torch.float32 else grad.float()
    return g, g.data_ptr()


# ----------------------------------------------------------------------------------------------------------------
# gae
# ----------------------------------------------------------------------------------------------------------------
def gae_(value, next_value, reward, done, traj_flag, gamma, lambda_, agents, mask_inplace=True):
    """value/next_value (T, C) contiguous fp32 CUDA; reward/done/traj (T, C/agents). next_value is masked in place."""
    T = value.shape[0]
    C = value.numel() // T if T > 0 else 0
    adv = torch.empty_like(value)
    if value.numel() == 0:
        return adv
    if C == 1 and agents == 1:
        # ONE sequence (the real PPO learner, ding/policy/ppo.py:280-282: n_sample steps of concatenated trajectories): the
        # segment-parallel single-CTA kernel of csrc/policy.cu instead of one lane walking all T steps
        return gae_returns_(value, next_value, reward, done, traj_flag, gamma, lambda_, 1, 0.0, False, False, mask_inplace)[0]
    with on_device(value.device):
        rc = lib().b200rl_gae(
            ptr(value), ptr(next_value), ptr(reward), ptr(done), ptr(traj_flag), ptr(adv), T, C, agents, float(gamma),
            float(lambda_), 1 if mask_inplace else 0, stream_ptr()
        )
    _lib.check(rc, 'b200rl_gae')
    return adv


def gae_returns_(value, next_value, reward, done, traj_flag, gamma, lambda_, agents, vscale, want_returns, want_stats,
                 mask_inplace=False, want_adv_stats=False):
    """gae + the pieces around it in PPOPolicy._forward_learn (ding/policy/ppo.py:274-297) -- csrc/policy.cu.
    -> (adv, unnormalized_return, value_out, return_out, stats3[, adv_stats2]) (None where not requested)."""
    T = value.shape[0]
    C = value.numel() // T
    dev = value.device
    adv = torch.empty_like(value)
    unnorm = torch.empty_like(value) if want_returns else None
    vout = torch.empty_like(value) if (want_returns and vscale != 0.0) else None
    rout = torch.empty_like(value) if (want_returns and vscale != 0.0) else None
    stats = torch.empty(3, dtype=torch.float32, device=dev) if want_stats else None
    astats = torch.empty(2, dtype=torch.float32, device=dev) if want_adv_stats else None
    with on_device(dev):
        ws = workspace(dev)
        rc = lib().b200rl_gae_returns(
            ptr(value), ptr(next_value), ptr(reward), ptr(done), ptr(traj_flag), T, C, agents, float(gamma), float(lambda_),
            1 if mask_inplace else 0, float(vscale), ptr(adv), ptr(unnorm), ptr(vout), ptr(rout), ptr(stats), ptr(astats),
            ptr(ws), ws.numel() * 4, stream_ptr()
        )
    _lib.check(rc, 'b200rl_gae_returns')
    if want_adv_stats:
        return adv, unnorm, vout, rout, stats, astats
    return adv, unnorm, vout, rout, stats


def adv_stats_(x):
    """{mean, std(unbiased) + 1e-8} of ``x`` as two device floats (ding/policy/ppo.py:304-306), one launch."""
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    with on_device(x.device):
        ws = workspace(x.device)
        rc = lib().b200rl_adv_stats(ptr(x), x.numel(), ptr(out), ptr(ws), ws.numel() * 4, stream_ptr())
    _lib.check(rc, 'b200rl_adv_stats')
    return out


def normalize_(x, stats):
    out = torch.empty_like(x)
    with on_device(x.device):
        rc = lib().b200rl_normalize(ptr(x), ptr(stats), x.numel(), ptr(out), stream_ptr())
    _lib.check(rc, 'b200rl_normalize')
    return out


# ----------------------------------------------------------------------------------------------------------------
# ppo
# ----------------------------------------------------------------------------------------------------------------
# The forward pass also writes the gradients for the upstream gradients it expects (the loss weights of the training
# loop, remembered on the device from the previous backward pass); backward() verifies them on the device and only
# recomputes on a mismatch -- exact for any upstream gradient, no host sync.  False: separate backward kernel always.
PPO_FUSED_BACKWARD = True
_PPO_HINT = {}


def ppo_hint(device, kind='ppo'):
    """Device-resident expectation of (d total/d policy_loss, d/d value_loss, d/d entropy_loss, d/d kl_div), one per call
    site kind: a policy-only caller (``ppo_policy_error``: no value term) must not disturb ``ppo_error``'s expectation."""
    key = (device.index, kind)
    h = _PPO_HINT.get(key)
    if h is None:
        init = [1.0, 0.5, -0.01, 0.0] if kind in ('ppo', 'happo') else [1.0, 0.0, -0.01, 0.0]
        h = torch.tensor(init, dtype=torch.float32, device=device)
        _PPO_HINT[key] = h
    return h


class PPOFunction(torch.autograd.Function):
    """Outputs: policy_loss, value_loss, entropy_loss, kl_div (differentiable 0-dim) and the raw 8-float result vector
    (non differentiable; [4]=approx_kl, [5]=clipfrac)."""

    @staticmethod
    def forward(ctx, logit_new, value_new, logit_old, action, value_old, adv, return_, weight, logit_pre, S, G, N,
                clip_ratio, use_value_clip, dual_clip, kl_type, hint_kind, adv_stats=None, factor=None):
        dev = logit_new.device
        ctx.hint_kind = hint_kind
        ctx.adv_stats = adv_stats  # {mean, std + 1e-8} device floats or None; kept alive for the backward launch
        ctx.factor = factor        # happo_error's per-sample factor (S,) or None; likewise
        out = torch.empty(8, dtype=torch.float32, device=dev)
        L = lib()
        tensors = (ptr(logit_new), ptr(logit_old), ptr(logit_pre), ptr(action), ptr(value_new), ptr(value_old),
                   ptr(adv), ptr(return_), ptr(weight))
        cfg = (S, G, N, clip_ratio, use_value_clip, dual_clip, kl_type, ptr(adv_stats), ptr(factor))
        ctx.fused = False
        want_grad = PPO_FUSED_BACKWARD and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        with on_device(dev):
            ws = workspace(dev)
            if want_grad:
                grad_logit = torch.empty_like(logit_new)
                grad_value = torch.empty_like(value_new)
                if L.b200rl_ppo_fused_supported(*tensors, ptr(grad_logit), G, N):
                    g_used = torch.empty(4, dtype=torch.float32, device=dev)
                    rc = L.b200rl_ppo_fwd_grad(*tensors, *cfg, ptr(ppo_hint(dev, hint_kind)), ptr(g_used), ptr(out),
                                               ptr(grad_logit), ptr(grad_value), ptr(ws), ws.numel() * 4,
                                               stream_ptr())
                    _lib.check(rc, 'b200rl_ppo_fwd_grad')
                    ctx.fused = True
                    ctx.spec = (grad_logit, grad_value, g_used)
            if not ctx.fused:
                rc = L.b200rl_ppo_fwd(*tensors, *cfg, ptr(out), ptr(ws), ws.numel() * 4, stream_ptr())
                _lib.check(rc, 'b200rl_ppo_fwd')
        ctx.save_for_backward(logit_new, value_new, logit_old, action, value_old, adv, return_, weight, logit_pre)
        ctx.cfg = cfg
        ctx.bwd_calls = 0
        ctx.mark_non_differentiable(out)
        p, v, e, k = out[0], out[1], out[2], out[3]
        return p, v, e, k, out

    @staticmethod
    def backward(ctx, g_p, g_v, g_e, g_k, _g_out):
        logit_new, value_new, logit_old, action, value_old, adv, return_, weight, logit_pre = ctx.saved_tensors
        dev = logit_new.device
        kp, pp = _g(g_p)
        kv, pv = _g(g_v)
        ke, pe = _g(g_e)
        kk, pk = _g(g_k)
        first = ctx.bwd_calls == 0
        ctx.bwd_calls += 1
        if ctx.fused and first:
            grad_logit, grad_value, g_used = ctx.spec  # valid if the expectation held; the kernel checks on the device
            ctx.spec = None  # sole owner now: autograd can adopt the buffers as .grad instead of cloning them
            p_used, p_hint = ptr(g_used), ptr(ppo_hint(dev, getattr(ctx, 'hint_kind', 'ppo')))
        else:  # no fused forward, or a repeated backward (the first call's buffers may now belong to .grad)
            grad_logit = torch.empty_like(logit_new)
            grad_value = torch.empty_like(value_new)
            p_used, p_hint = None, None
        with on_device(dev):
            rc = lib().b200rl_ppo_bwd(
                ptr(logit_new), ptr(logit_old), ptr(logit_pre), ptr(action), ptr(value_new), ptr(value_old), ptr(adv),
                ptr(return_), ptr(weight), *ctx.cfg, pp, pv, pe, pk, p_used, p_hint, ptr(grad_logit), ptr(grad_value),
                stream_ptr()
            )
        _lib.check(rc, 'b200rl_ppo_bwd')
        return (grad_logit, grad_value) + (None, ) * 17


class GAEPPOFunction(torch.autograd.Function):
    """gae -> ppo_error in one launch (csrc/fused.cu).  Outputs: adv (T,B; non differentiable), the four losses and the
    raw result vector; backward is the same device-verified scheme as PPOFunction."""

    @staticmethod
    def forward(ctx, logit_new, value_new, value, next_value, reward, done, traj_flag, logit_old, action, value_old,
                return_, weight, logit_pre, T, B, N, gamma, lambda_, clip_ratio, use_value_clip, dual_clip, kl_type):
        dev = logit_new.device
        out = torch.empty(8, dtype=torch.float32, device=dev)
        adv = torch.empty_like(value)
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        grad_logit = torch.empty_like(logit_new) if want_grad else None
        grad_value = torch.empty_like(value_new) if want_grad else None
        g_used = torch.empty(4, dtype=torch.float32, device=dev) if want_grad else None
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_gae_ppo_fwd_grad(
                ptr(value), ptr(next_value), ptr(reward), ptr(done), ptr(traj_flag), T, B, gamma, lambda_, 1,
                ptr(logit_new), ptr(logit_old), ptr(logit_pre), ptr(action), ptr(value_new), ptr(value_old),
                ptr(return_), ptr(weight), N, clip_ratio, use_value_clip, dual_clip, kl_type,
                ptr(ppo_hint(dev)) if want_grad else None, ptr(g_used), ptr(adv), ptr(out), ptr(grad_logit),
                ptr(grad_value), ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_gae_ppo_fwd_grad')
        ctx.save_for_backward(logit_new, value_new, logit_old, action, value_old, adv, return_, weight, logit_pre)
        ctx.cfg = (T * B, 1, N, clip_ratio, use_value_clip, dual_clip, kl_type, None, None)  # no adv_stats, no factor
        ctx.fused = want_grad
        ctx.spec = (grad_logit, grad_value, g_used)
        ctx.bwd_calls = 0
        ctx.mark_non_differentiable(out, adv)
        return adv, out[0], out[1], out[2], out[3], out

    @staticmethod
    def backward(ctx, _g_adv, g_p, g_v, g_e, g_k, _g_out):
        grads = PPOFunction.backward(ctx, g_p, g_v, g_e, g_k, None)
        return (grads[0], grads[1]) + (None, ) * 20


def ppo_value_(value_new, value_old, return_, weight, clip_ratio, use_value_clip):
    """ppo_value_error (ppo.py:233-275): loss and, if needed, its gradient for a unit upstream gradient in one launch."""
    dev = value_new.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    want_grad = value_new.requires_grad and torch.is_grad_enabled()
    dvalue = torch.empty_like(value_new) if want_grad else None
    with on_device(dev):
        ws = workspace(dev)
        rc = lib().b200rl_ppo_value_fwd(
            ptr(value_new), ptr(value_old), ptr(return_), ptr(weight), value_new.numel(), float(clip_ratio),
            1 if use_value_clip else 0, ptr(loss), ptr(dvalue), ptr(ws), ws.numel() * 4, stream_ptr()
        )
    _lib.check(rc, 'b200rl_ppo_value_fwd')
    if want_grad:
        return _ScaleSaved.apply(value_new, loss, dvalue)
    return loss


def ppg_bc_(logit_new, logit_old, action):
    """behavioural-cloning term of ppg_joint_error (ppg.py:62-67): value (NaN-propagating, as the reference) and gradient."""
    dev = logit_new.device
    B, N = logit_new.shape
    loss = torch.empty((), dtype=torch.float32, device=dev)
    want_grad = logit_new.requires_grad and torch.is_grad_enabled()
    dlogit = torch.empty_like(logit_new) if want_grad else None
    with on_device(dev):
        ws = workspace(dev)
        rc = lib().b200rl_ppg_bc_fwd(ptr(logit_new), ptr(logit_old), ptr(action), B, N, ptr(loss), ptr(dlogit), ptr(ws),
                                     ws.numel() * 4, stream_ptr())
    _lib.check(rc, 'b200rl_ppg_bc_fwd')
    if want_grad:
        return _ScaleSaved.apply(logit_new, loss, dlogit)
    return loss


# ----------------------------------------------------------------------------------------------------------------
# q n-step TD
# ----------------------------------------------------------------------------------------------------------------
class QNStepTDFunction(torch.autograd.Function):
    """Outputs: loss and td_error_per_sample (both differentiable w.r.t. q, like the reference's, td.py:718-719), the
    detached n-step target and (sequence form) the priority mix (non differentiable).

    ONE forward launch also writes d loss / d q for a unit upstream gradient; ``backward`` hands that buffer to autograd
    after a verification launch that returns at once when the upstream gradient really was 1 (and no gradient arrived
    through td_error_per_sample), and recomputes otherwise -- exact for any upstream gradient, no host sync."""

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, weight, value_gamma, vg_stride, gamma_ps, S, G,
                N, nstep, gamma, cum_reward, rescale, eps, criterion, crit_param, group_mean, seq_len, priority_mix,
                want_priority):
        dev = q.device
        R = S * G
        loss = torch.empty((), dtype=torch.float32, device=dev)
        td = torch.empty(S if group_mean else R, dtype=torch.float32, device=dev)
        dcrit = torch.empty(R, dtype=torch.float32, device=dev)
        target = torch.empty(R, dtype=torch.float32, device=dev)
        prio = torch.empty(S // seq_len, dtype=torch.float32, device=dev) if want_priority else None
        grad_unit = torch.empty(R, N, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_qntd_fwd(
                ptr(q), ptr(next_n_q), ptr(action), ptr(next_n_action), ptr(reward), ptr(done), ptr(weight),
                ptr(value_gamma), vg_stride, ptr(gamma_ps), S, G, N, nstep, gamma, cum_reward, rescale, eps, criterion,
                crit_param, group_mean, seq_len, priority_mix, ptr(loss), ptr(td), ptr(dcrit), ptr(target),
                ptr(grad_unit), ptr(prio), ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_qntd_fwd')
        ctx.save_for_backward(dcrit, action, weight)
        ctx.cfg = (S, G, N, group_mean, seq_len)
        ctx.q_shape = q.shape
        ctx.spec = grad_unit
        ctx.set_materialize_grads(False)
        if prio is None:
            prio = torch.empty(0, dtype=torch.float32, device=dev)
        ctx.mark_non_differentiable(target, prio)
        return loss, td, target, prio

    @staticmethod
    def backward(ctx, g_loss, g_td, _g_target, _g_prio):
        if g_loss is None and g_td is None:
            return (None, ) * 24
        dcrit, action, weight = ctx.saved_tensors
        S, G, N, group_mean, seq_len = ctx.cfg
        keep, pg = _g(g_loss)
        keep_td, ptd = None, None
        if g_td is not None:
            keep_td = f32c(g_td)
            ptd = keep_td.data_ptr()
        grad_q, skip = ctx.spec, 0
        ctx.spec = None  # sole owner now: autograd can adopt the buffer as .grad instead of cloning it
        if grad_q is not None and g_td is None:
            skip = 1
        else:
            grad_q = torch.empty(S * G, N, dtype=torch.float32, device=dcrit.device)
        with on_device(dcrit.device):
            rc = lib().b200rl_qntd_bwd(ptr(dcrit), ptr(weight), ptr(action), pg, ptd, S, G, N, group_mean, seq_len, skip,
                                       ptr(grad_q), stream_ptr())
        _lib.check(rc, 'b200rl_qntd_bwd')
        return (grad_q.view(ctx.q_shape), ) + (None, ) * 23


# ----------------------------------------------------------------------------------------------------------------
# ACER heads (csrc/acer.cu): un-reduced per-transition losses, plain forward / backward launches
# ----------------------------------------------------------------------------------------------------------------
class AcerPolicyFunction(torch.autograd.Function):
    """(actor_loss, bias_correction_loss), each (M,); differentiable w.r.t. target_logit only (acer.py:44-56)."""

    @staticmethod
    def forward(ctx, target_logit, q_values, q_retraces, v_pred, actions, ratio, M, N, c_clip_ratio):
        dev = target_logit.device
        actor = torch.empty(M, dtype=torch.float32, device=dev)
        bias = torch.empty(M, dtype=torch.float32, device=dev)
        with on_device(dev):
            rc = lib().b200rl_acer_policy_fwd(ptr(q_values), ptr(q_retraces), ptr(v_pred), ptr(target_logit), ptr(actions),
                                              ptr(ratio), M, N, c_clip_ratio, ptr(actor), ptr(bias), stream_ptr())
        _lib.check(rc, 'b200rl_acer_policy_fwd')
        ctx.save_for_backward(target_logit, q_values, q_retraces, v_pred, actions, ratio)
        ctx.cfg = (M, N, c_clip_ratio)
        ctx.set_materialize_grads(False)
        return actor, bias

    @staticmethod
    def backward(ctx, g_actor, g_bias):
        if g_actor is None and g_bias is None:
            return (None, ) * 9
        target_logit, q_values, q_retraces, v_pred, actions, ratio = ctx.saved_tensors
        M, N, c = ctx.cfg
        ga = f32c(g_actor) if g_actor is not None else None
        gb = f32c(g_bias) if g_bias is not None else None
        grad = torch.empty_like(target_logit)
        with on_device(target_logit.device):
            rc = lib().b200rl_acer_policy_bwd(ptr(q_values), ptr(q_retraces), ptr(v_pred), ptr(target_logit), ptr(actions),
                                              ptr(ratio), ptr(ga), ptr(gb), M, N, c, ptr(grad), stream_ptr())
        _lib.check(rc, 'b200rl_acer_policy_bwd')
        return (grad, ) + (None, ) * 8


class AcerValueFunction(torch.autograd.Function):
    """critic_loss (M,) = 0.5 (q_retraces - q_values[a])^2; differentiable w.r.t. q_values (acer.py:81-82)."""

    @staticmethod
    def forward(ctx, q_values, q_retraces, actions, M, N):
        loss = torch.empty(M, dtype=torch.float32, device=q_values.device)
        with on_device(q_values.device):
            rc = lib().b200rl_acer_value_fwd(ptr(q_values), ptr(q_retraces), ptr(actions), M, N, ptr(loss), stream_ptr())
        _lib.check(rc, 'b200rl_acer_value_fwd')
        ctx.save_for_backward(q_values, q_retraces, actions)
        ctx.cfg = (M, N)
        return loss

    @staticmethod
    def backward(ctx, g):
        q_values, q_retraces, actions = ctx.saved_tensors
        M, N = ctx.cfg
        gg = f32c(g)
        grad = torch.empty_like(q_values)
        with on_device(q_values.device):
            rc = lib().b200rl_acer_value_bwd(ptr(q_values), ptr(q_retraces), ptr(actions), ptr(gg), M, N, ptr(grad),
                                             stream_ptr())
        _lib.check(rc, 'b200rl_acer_value_bwd')
        return grad, None, None, None, None


def acer_trust_region_(grad, avg_logit, delta):
    N = grad.shape[-1]
    M = grad.numel() // N
    out = torch.empty_like(grad)
    with on_device(grad.device):
        rc = lib().b200rl_acer_trust_region(ptr(grad), ptr(avg_logit), M, N, float(delta), ptr(out), stream_ptr())
    _lib.check(rc, 'b200rl_acer_trust_region')
    return out


# ----------------------------------------------------------------------------------------------------------------
# quantile-regression n-step TD (QR-DQN / IQN / FQF)
# ----------------------------------------------------------------------------------------------------------------
class QuantileTDFunction(torch.autograd.Function):
    """loss and the per-sample losses (both differentiable w.r.t. q, as in the reference, td.py:1166,:1346,:1436); the forward
    launch also writes d loss / d q for a unit upstream gradient (verified on the device by the backward launch, as
    QNStepTDFunction).  ``q_strides`` / ``nq_strides`` / ``tau_strides``: element strides of (sample, quantile[, action])."""

    @staticmethod
    def forward(ctx, q, next_n_q, action, next_n_action, reward, done, tau, weight, value_gamma, vg_stride, B, N, n_tau,
                n_tau_prime, nstep, gamma, q_strides, nq_strides, tau_strides, form, kappa):
        dev = q.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        td = torch.empty(B, dtype=torch.float32, device=dev)
        dtheta = torch.empty(B, n_tau, dtype=torch.float32, device=dev)
        grad_unit = torch.empty_like(q) if ctx.needs_input_grad[0] else None
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_quantile_td_fwd(
                ptr(q), ptr(next_n_q), ptr(action), ptr(next_n_action), ptr(reward), ptr(done), ptr(tau), ptr(weight),
                ptr(value_gamma), vg_stride, B, N, n_tau, n_tau_prime, nstep, gamma, *q_strides, *nq_strides, *tau_strides,
                form, kappa, ptr(loss), ptr(td), ptr(dtheta), ptr(grad_unit), ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_quantile_td_fwd')
        ctx.save_for_backward(dtheta, action, weight)
        ctx.cfg = (B, N, n_tau, q_strides)
        ctx.q_shape = q.shape
        ctx.spec = grad_unit
        ctx.set_materialize_grads(False)
        return loss, td

    @staticmethod
    def backward(ctx, g_loss, g_td):
        if g_loss is None and g_td is None:
            return (None, ) * 21
        dtheta, action, weight = ctx.saved_tensors
        B, N, n_tau, q_strides = ctx.cfg
        keep, pg = _g(g_loss)
        keep_td = f32c(g_td) if g_td is not None else None
        grad_q, skip = ctx.spec, 1
        ctx.spec = None
        if grad_q is None or g_td is not None:  # repeated backward / a gradient through the per-sample losses
            grad_q, skip = torch.empty(ctx.q_shape, dtype=torch.float32, device=dtheta.device), 0
        with on_device(dtheta.device):
            rc = lib().b200rl_quantile_td_bwd(ptr(dtheta), ptr(weight), ptr(action), pg, ptr(keep_td), B, N, n_tau,
                                              *q_strides, skip, ptr(grad_q), stream_ptr())
        _lib.check(rc, 'b200rl_quantile_td_bwd')
        return (grad_q, ) + (None, ) * 20


# ----------------------------------------------------------------------------------------------------------------
# distributional n-step TD (C51)
# ----------------------------------------------------------------------------------------------------------------
class DistNStepTDFunction(torch.autograd.Function):
    """loss (differentiable w.r.t. dist) and the unweighted per-sample error; the forward launch also writes the gradient for
    a unit upstream gradient (verified on the device by the backward launch, as QNStepTDFunction)."""

    @staticmethod
    def forward(ctx, dist, next_n_dist, act, next_n_act, reward, done, weight, w_stride, value_gamma, vg_stride,
                support, B, A, N, n_atom, nstep, gamma, v_min, v_max, bad_flag):
        dev = dist.device
        R = B * A
        loss = torch.empty((), dtype=torch.float32, device=dev)
        td = torch.empty(R, dtype=torch.float32, device=dev)
        proj = torch.empty(R, n_atom, dtype=torch.float32, device=dev)
        grad_unit = torch.empty_like(dist) if ctx.needs_input_grad[0] else None
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_dntd_fwd(
                ptr(dist), ptr(next_n_dist), ptr(act), ptr(next_n_act), ptr(reward), ptr(done), ptr(weight), w_stride,
                ptr(value_gamma), vg_stride, ptr(support), B, A, N, n_atom, nstep, gamma, v_min, v_max, ptr(loss),
                ptr(td), ptr(proj), ptr(bad_flag), ptr(grad_unit), ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_dntd_fwd')
        ctx.save_for_backward(dist, act, proj, weight)
        ctx.cfg = (R, N, n_atom, w_stride)
        ctx.spec = grad_unit
        ctx.set_materialize_grads(False)
        return loss, td

    @staticmethod
    def backward(ctx, g_loss, g_td):
        if g_loss is None and g_td is None:
            return (None, ) * 20
        dist, act, proj, weight = ctx.saved_tensors
        R, N, n_atom, w_stride = ctx.cfg
        keep, pg = _g(g_loss)
        keep_td = f32c(g_td) if g_td is not None else None
        grad, skip = ctx.spec, 1
        ctx.spec = None
        if grad is None or g_td is not None:  # a repeated backward (the first buffer may belong to .grad) / per-sample path
            grad, skip = torch.empty_like(dist), 0
        with on_device(dist.device):
            rc = lib().b200rl_dntd_bwd(
                ptr(dist), ptr(act), ptr(proj), ptr(weight), w_stride, pg, ptr(keep_td), R, N, n_atom, skip, ptr(grad),
                stream_ptr()
            )
        _lib.check(rc, 'b200rl_dntd_bwd')
        return (grad, ) + (None, ) * 19


# ----------------------------------------------------------------------------------------------------------------
# lambda returns / TD(lambda)
# ----------------------------------------------------------------------------------------------------------------
def lambda_returns_(value, reward, gammas, gamma, lambdas, lambda_, done, upgo_mode):
    T, B = reward.shape
    ret = torch.empty_like(reward)
    with on_device(value.device):
        rc = lib().b200rl_lambda_returns(
            ptr(value), ptr(reward), ptr(gammas), float(gamma), ptr(lambdas), float(lambda_), ptr(done),
            1 if upgo_mode else 0, T, B, ptr(ret), stream_ptr()
        )
    _lib.check(rc, 'b200rl_lambda_returns')
    return ret


class LambdaReturnsFunction(torch.autograd.Function):
    """generalized_lambda_returns / upgo_returns with the reference's differentiability (td.py:1574-1651 is plain torch
    arithmetic): gradients reach bootstrap_values, rewards and -- when they are tensors that require grad -- gammas and
    lambda_.  Backward is the transposed scan (csrc/td.cu: lambda_returns_bwd_kernel), one launch."""

    @staticmethod
    def forward(ctx, value, reward, gammas, lambdas, done, gamma, lambda_, upgo_mode):
        ret = lambda_returns_(value, reward, gammas, gamma, lambdas, lambda_, done, upgo_mode)
        ctx.save_for_backward(value, reward, gammas, lambdas, done, ret)
        ctx.scal = (float(gamma), float(lambda_), 1 if upgo_mode else 0)
        return ret

    @staticmethod
    def backward(ctx, g_ret):
        value, reward, gammas, lambdas, done, ret = ctx.saved_tensors
        gamma, lambda_, upgo = ctx.scal
        T, B = reward.shape
        g = f32c(g_ret)
        need = ctx.needs_input_grad
        gv = torch.empty_like(value)
        gr = torch.empty_like(reward) if need[1] else None
        gg = torch.empty_like(reward) if (need[2] and gammas is not None) else None
        gl = torch.empty_like(reward) if (need[3] and lambdas is not None) else None
        with on_device(value.device):
            rc = lib().b200rl_lambda_returns_bwd(
                ptr(g), ptr(value), ptr(reward), ptr(ret), ptr(gammas), gamma, ptr(lambdas), lambda_, ptr(done), upgo,
                T, B, ptr(gv), ptr(gr), ptr(gg), ptr(gl), stream_ptr()
            )
        _lib.check(rc, 'b200rl_lambda_returns_bwd')
        return gv if need[0] else None, gr, gg, gl, None, None, None, None


class TBCrossEntropyFunction(torch.autograd.Function):
    """tb_cross_entropy (upgo.py:7-43): ce (TB) = sum_k mask_k * log softmax(logit)[label]; gradient reaches ``logit``."""

    @staticmethod
    def forward(ctx, logit, label, mask, TB, K, N):
        ce = torch.empty(TB, dtype=torch.float32, device=logit.device)
        with on_device(logit.device):
            rc = lib().b200rl_tb_cross_entropy_fwd(ptr(logit), ptr(label), ptr(mask), TB, K, N, ptr(ce), stream_ptr())
        _lib.check(rc, 'b200rl_tb_cross_entropy_fwd')
        ctx.save_for_backward(logit, label, mask)
        ctx.cfg = (TB, K, N)
        return ce

    @staticmethod
    def backward(ctx, g_ce):
        logit, label, mask = ctx.saved_tensors
        TB, K, N = ctx.cfg
        g = f32c(g_ce).reshape(-1)
        grad = torch.empty_like(logit)
        with on_device(logit.device):
            rc = lib().b200rl_tb_cross_entropy_bwd(ptr(logit), ptr(label), ptr(mask), ptr(g), TB, K, N, ptr(grad),
                                                   stream_ptr())
        _lib.check(rc, 'b200rl_tb_cross_entropy_bwd')
        return grad, None, None, None, None, None


class _ScaleSaved(torch.autograd.Function):
    """loss whose gradient w.r.t. ``x`` was produced by the forward kernel for a unit upstream gradient."""

    @staticmethod
    def forward(ctx, x, loss, saved_grad):
        ctx.save_for_backward(saved_grad)
        return loss.view_as(loss)

    @staticmethod
    def backward(ctx, g):
        saved, = ctx.saved_tensors
        out = torch.empty_like(saved)
        keep, pg = _g(g)
        with on_device(saved.device):
            rc = lib().b200rl_scale(pg, ptr(saved), ptr(out), saved.numel(), stream_ptr())
        _lib.check(rc, 'b200rl_scale')
        return out, None, None


def td_lambda_(value, reward, weight, gamma, lambda_):
    T, B = reward.shape
    dev = value.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dvalue = torch.empty_like(value)
    with on_device(dev):
        ws = workspace(dev)
        rc = lib().b200rl_td_lambda_fwd(
            ptr(value), ptr(reward), ptr(weight), float(gamma), float(lambda_), T, B, ptr(loss), ptr(dvalue), ptr(ws),
            ws.numel() * 4, stream_ptr()
        )
    _lib.check(rc, 'b200rl_td_lambda_fwd')
    if value.requires_grad and torch.is_grad_enabled():
        return _ScaleSaved.apply(value, loss, dvalue)
    return loss


# ----------------------------------------------------------------------------------------------------------------
# UPGO head
# ----------------------------------------------------------------------------------------------------------------
class UPGOFunction(torch.autograd.Function):
    """upgo_loss head (upgo.py:77-111).  The forward launch also writes d loss / d logit for a unit upstream gradient while each
    row is still in L1 (ONE pass over the logits); backward verifies the upstream gradient on the device and recomputes only
    when it is not 1 (or on a repeated backward)."""

    @staticmethod
    def forward(ctx, logit, action, mask, rho, ret, value, TB, K, N):
        dev = logit.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        adv = torch.empty(TB, dtype=torch.float32, device=dev)
        grad_unit = torch.empty_like(logit) if ctx.needs_input_grad[0] else None
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_upgo_head_fwd(
                ptr(logit), ptr(action), ptr(mask), ptr(rho), ptr(ret), ptr(value), TB, K, N, ptr(loss), ptr(adv),
                ptr(grad_unit), ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_upgo_head_fwd')
        ctx.save_for_backward(logit, action, mask, adv)
        ctx.cfg = (TB, K, N)
        ctx.spec = grad_unit
        return loss

    @staticmethod
    def backward(ctx, g):
        logit, action, mask, adv = ctx.saved_tensors
        TB, K, N = ctx.cfg
        grad, skip = ctx.spec, 1
        ctx.spec = None  # sole owner now: autograd can adopt the buffer as .grad
        if grad is None:  # a repeated backward: the first buffer may belong to .grad
            grad, skip = torch.empty_like(logit), 0
        keep, pg = _g(g)
        with on_device(logit.device):
            rc = lib().b200rl_upgo_head_bwd(
                ptr(logit), ptr(action), ptr(mask), ptr(adv), pg, TB, K, N, skip, ptr(grad), stream_ptr()
            )
        _lib.check(rc, 'b200rl_upgo_head_bwd')
        return (grad, ) + (None, ) * 8


_HEAD_HINT = {}


def head_hint(device, kind, init):
    """device-resident expectation of the upstream gradients of a loss head (one record per head kind and device)"""
    key = (device.index, kind)
    h = _HEAD_HINT.get(key)
    if h is None:
        h = torch.tensor(init, dtype=torch.float32, device=device)
        _HEAD_HINT[key] = h
    return h


class A2CFunction(torch.autograd.Function):
    """a2c_error (ding/rl_utils/a2c.py:10-44): three differentiable 0-dim losses; gradients reach logit and value.  One launch
    forward (losses + gradients for the expected upstream gradients), one verification launch backward (csrc/heads.cu)."""

    @staticmethod
    def forward(ctx, logit, value, action, adv, return_, weight, S, N):
        dev = logit.device
        out = torch.empty(4, dtype=torch.float32, device=dev)
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gl = torch.empty_like(logit) if want else None
        gv = torch.empty_like(value) if want else None
        g_used = torch.empty(4, dtype=torch.float32, device=dev) if want else None
        ctx.hint = head_hint(dev, 'a2c', [1.0, 0.5, -0.01, 0.0])
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_a2c_fwd_grad(ptr(logit), ptr(action), ptr(value), ptr(adv), ptr(return_), ptr(weight), S, N,
                                           ptr(ctx.hint) if want else None, 0, None, None, None, ptr(g_used), None, ptr(out),
                                           ptr(gl), ptr(gv), ptr(ws), ws.numel() * 4, stream_ptr())
        _lib.check(rc, 'b200rl_a2c_fwd_grad')
        ctx.save_for_backward(logit, value, action, adv, return_, weight)
        ctx.cfg = (S, N)
        ctx.spec = (gl, gv, g_used)
        ctx.set_materialize_grads(False)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_p, g_v, g_e):
        logit, value, action, adv, return_, weight = ctx.saved_tensors
        S, N = ctx.cfg
        kp, pp = _g(g_p)
        kv, pv = _g(g_v)
        ke, pe = _g(g_e)
        spec, ctx.spec = ctx.spec, None
        if spec is not None and spec[0] is not None:
            gl, gv, g_used = spec
        else:  # a repeated backward: the first call's buffers may now belong to .grad -> recompute into fresh ones
            gl, gv = torch.empty_like(logit), torch.empty_like(value)
            g_used = torch.full((4, ), float('nan'), dtype=torch.float32, device=logit.device)
        with on_device(logit.device):
            ws = workspace(logit.device)
            rc = lib().b200rl_a2c_fwd_grad(ptr(logit), ptr(action), ptr(value), ptr(adv), ptr(return_), ptr(weight), S, N,
                                           None, 1, pp, pv, pe, ptr(g_used), ptr(ctx.hint), None, ptr(gl), ptr(gv), ptr(ws),
                                           ws.numel() * 4, stream_ptr())
        _lib.check(rc, 'b200rl_a2c_fwd_grad(verify)')
        return gl, gv, None, None, None, None, None, None


class PPOContinuousFunction(torch.autograd.Function):
    """ppo_error_continuous (ding/rl_utils/ppo.py:278-374): gradients reach mu_new, sigma_new and value_new (csrc/heads.cu)."""

    @staticmethod
    def forward(ctx, mu, sigma, value_new, mu_old, sigma_old, mu_pre, sigma_pre, action, value_old, adv, return_, weight, S,
                D, clip_ratio, use_value_clip, dual_clip, kl_type, factor=None):
        dev = mu.device
        ctx.factor = factor  # happo_error_continuous's per-sample factor (S,) or None; kept alive for the backward launch
        out = torch.empty(8, dtype=torch.float32, device=dev)
        want = any(ctx.needs_input_grad[:3])
        gm = torch.empty_like(mu) if want else None
        gs = torch.empty_like(sigma) if want else None
        gv = torch.empty_like(value_new) if want else None
        g_used = torch.empty(4, dtype=torch.float32, device=dev) if want else None
        ctx.hint = head_hint(dev, 'ppoc' if factor is None else 'happoc', [1.0, 0.5, -0.01, 0.0])
        ctx.args = (ptr(factor), S, D, clip_ratio, use_value_clip, dual_clip, kl_type)
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_ppo_continuous_fwd_grad(
                ptr(mu), ptr(sigma), ptr(mu_old), ptr(sigma_old), ptr(mu_pre), ptr(sigma_pre), ptr(action), ptr(value_new),
                ptr(value_old), ptr(adv), ptr(return_), ptr(weight), *ctx.args, ptr(ctx.hint) if want else None, 0, None,
                None, None, None, ptr(g_used), None, ptr(out), ptr(gm), ptr(gs), ptr(gv), ptr(ws), ws.numel() * 4,
                stream_ptr())
        _lib.check(rc, 'b200rl_ppo_continuous_fwd_grad')
        ctx.save_for_backward(mu, sigma, value_new, mu_old, sigma_old, mu_pre, sigma_pre, action, value_old, adv, return_,
                              weight)
        ctx.spec = (gm, gs, gv, g_used)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(out)
        return out[0], out[1], out[2], out[3], out

    @staticmethod
    def backward(ctx, g_p, g_v, g_e, g_k, _g_out):
        (mu, sigma, value_new, mu_old, sigma_old, mu_pre, sigma_pre, action, value_old, adv, return_,
         weight) = ctx.saved_tensors
        keep = [_g(x) for x in (g_p, g_v, g_e, g_k)]
        spec, ctx.spec = ctx.spec, None
        if spec is not None and spec[0] is not None:
            gm, gs, gv, g_used = spec
        else:
            gm, gs, gv = torch.empty_like(mu), torch.empty_like(sigma), torch.empty_like(value_new)
            g_used = torch.full((4, ), float('nan'), dtype=torch.float32, device=mu.device)
        with on_device(mu.device):
            ws = workspace(mu.device)
            rc = lib().b200rl_ppo_continuous_fwd_grad(
                ptr(mu), ptr(sigma), ptr(mu_old), ptr(sigma_old), ptr(mu_pre), ptr(sigma_pre), ptr(action), ptr(value_new),
                ptr(value_old), ptr(adv), ptr(return_), ptr(weight), *ctx.args, None, 1, keep[0][1], keep[1][1], keep[2][1],
                keep[3][1], ptr(g_used), ptr(ctx.hint), None, ptr(gm), ptr(gs), ptr(gv), ptr(ws), ws.numel() * 4,
                stream_ptr())
        _lib.check(rc, 'b200rl_ppo_continuous_fwd_grad(verify)')
        return (gm, gs, gv) + (None, ) * 16


class ImpalaMaskFunction(torch.autograd.Function):
    """IMPALAPolicy._reshape_data masking (ding/policy/impala.py:316-322): values (T+1, B) (differentiable), rewards, done
    (T, B) -> (values', rewards', weights).  The reference multiplies ``values[1:]`` in place, so gradient reaches the critic
    output through the mask; backward is the same kernel applied to the upstream gradient."""

    @staticmethod
    def forward(ctx, values, rewards, done):
        T, B = rewards.shape
        vo = torch.empty_like(values)
        ro = torch.empty_like(rewards)
        wo = torch.empty_like(rewards)
        with on_device(values.device):
            rc = lib().b200rl_impala_mask(ptr(values), ptr(rewards), ptr(done), T, B, ptr(vo), ptr(ro), ptr(wo), stream_ptr())
        _lib.check(rc, 'b200rl_impala_mask')
        ctx.save_for_backward(done)
        ctx.mark_non_differentiable(ro, wo)
        return vo, ro, wo

    @staticmethod
    def backward(ctx, g_v, _g_r, _g_w):
        done, = ctx.saved_tensors
        T, B = done.shape
        g = f32c(g_v)
        out = torch.empty_like(g)
        with on_device(g.device):
            rc = lib().b200rl_impala_mask(ptr(g), None, ptr(done), T, B, ptr(out), None, None, stream_ptr())
        _lib.check(rc, 'b200rl_impala_mask')
        return out, None, None


# ----------------------------------------------------------------------------------------------------------------
# V-trace
# ----------------------------------------------------------------------------------------------------------------
# The one-launch kernel (csrc/vtws.cu) writes the gradients in the forward pass for the upstream gradients it expects (the
# loss weights of the training loop, remembered on the device from the previous backward pass; IMPALA's defaults to start
# with); backward() verifies them on the device and only recomputes on a mismatch.  False: rows / scan / backward kernels.
VTRACE_FUSED = True
_VT_HINT = {}


def vtrace_hint(device):
    """Device-resident expectation of (d total/d policy_loss, d/d value_loss, d/d entropy_loss)."""
    h = _VT_HINT.get(device.index)
    if h is None:
        h = torch.tensor([1.0, 0.5, -0.01], dtype=torch.float32, device=device)
        _VT_HINT[device.index] = h
    return h


class VTraceFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, target_output, value, behaviour_output, action, reward, weight, gamma, lambda_, rho_clip, c_clip,
                rho_pg_clip):
        T, B = reward.shape
        N = target_output.shape[-1]
        dev = target_output.device
        out = torch.empty(4, dtype=torch.float32, device=dev)
        L = lib()
        tensors = (ptr(target_output), ptr(behaviour_output), ptr(action), ptr(value), ptr(reward), ptr(weight))
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        ctx.cfg = (T, B, N)
        ctx.fused = False
        ctx.bwd_calls = 0
        with on_device(dev):
            ws = workspace(dev)
            if VTRACE_FUSED:
                grad_logit = torch.empty_like(target_output) if want_grad else None
                grad_value = torch.empty(T + 1, B, dtype=torch.float32, device=dev) if want_grad else None
                if L.b200rl_vtrace_fused_supported(*tensors, T, B, N, ptr(grad_logit), ptr(grad_value)):
                    g_used = torch.empty(3, dtype=torch.float32, device=dev) if want_grad else None
                    rc = L.b200rl_vtrace_fwd_grad(
                        *tensors, T, B, N, gamma, lambda_, rho_clip, c_clip, rho_pg_clip,
                        ptr(vtrace_hint(dev)) if want_grad else None, 0, None, None, None, ptr(g_used), None, ptr(out),
                        ptr(grad_logit), ptr(grad_value), ptr(ws), ws.numel() * 4, stream_ptr()
                    )
                    _lib.check(rc, 'b200rl_vtrace_fwd_grad')
                    ctx.fused = True
                    ctx.spec = (grad_logit, grad_value, g_used)
                    ctx.scal = (gamma, lambda_, rho_clip, c_clip, rho_pg_clip)
                    ctx.save_for_backward(target_output, behaviour_output, action, value, reward, weight)
                    return out[0], out[1], out[2]
            lp = torch.empty(T, B, dtype=torch.float32, device=dev)
            cpg = torch.empty(T, B, dtype=torch.float32, device=dev)
            dv = torch.empty(T, B, dtype=torch.float32, device=dev)
            rc = L.b200rl_vtrace_fwd(
                *tensors, T, B, N, gamma, lambda_, rho_clip, c_clip, rho_pg_clip, ptr(out), ptr(lp), ptr(cpg), ptr(dv),
                ptr(ws), ws.numel() * 4, stream_ptr()
            )
        _lib.check(rc, 'b200rl_vtrace_fwd')
        ctx.save_for_backward(target_output, action, weight, cpg, dv)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_p, g_v, g_e):
        T, B, N = ctx.cfg
        kp, pp = _g(g_p)
        kv, pv = _g(g_v)
        ke, pe = _g(g_e)
        if ctx.fused:
            target_output, behaviour_output, action, value, reward, weight = ctx.saved_tensors
            dev = target_output.device
            first = ctx.bwd_calls == 0
            ctx.bwd_calls += 1
            if first:
                grad_logit, grad_value, g_used = ctx.spec  # valid if the expectation held; the kernel checks on the device
                ctx.spec = None  # sole owner now: autograd can adopt the buffers as .grad instead of cloning them
            else:  # a repeated backward: the first call's buffers may now belong to .grad -> recompute into fresh ones
                grad_logit = torch.empty_like(target_output)
                grad_value = torch.empty(T + 1, B, dtype=torch.float32, device=dev)
                g_used = torch.full((3, ), float('nan'), dtype=torch.float32, device=dev)
            with on_device(dev):
                ws = workspace(dev)
                rc = lib().b200rl_vtrace_fwd_grad(
                    ptr(target_output), ptr(behaviour_output), ptr(action), ptr(value), ptr(reward), ptr(weight), T, B,
                    N, *ctx.scal, None, 1, pp, pv, pe, ptr(g_used), ptr(vtrace_hint(dev)), None, ptr(grad_logit),
                    ptr(grad_value), ptr(ws), ws.numel() * 4, stream_ptr()
                )
            _lib.check(rc, 'b200rl_vtrace_fwd_grad(verify)')
            return (grad_logit, grad_value) + (None, ) * 9
        target_output, action, weight, cpg, dv = ctx.saved_tensors
        dev = target_output.device
        grad_logit = torch.empty_like(target_output)
        grad_value = torch.empty(T + 1, B, dtype=torch.float32, device=dev)
        with on_device(dev):
            rc = lib().b200rl_vtrace_bwd(
                ptr(target_output), ptr(action), ptr(weight), ptr(cpg), ptr(dv), pp, pv, pe, T, B, N, ptr(grad_logit),
                ptr(grad_value), stream_ptr()
            )
        _lib.check(rc, 'b200rl_vtrace_bwd')
        return (grad_logit, grad_value) + (None, ) * 9


class VTraceContinuousFunction(torch.autograd.Function):
    """vtrace_error_continuous_action (ding/rl_utils/vtrace.py:139-212): rows kernel -> shared scan -> backward rows kernel."""

    @staticmethod
    def forward(ctx, mu, sigma, value, mu_b, sigma_b, action, reward, weight, D, gamma, lambda_, rho_clip, c_clip,
                rho_pg_clip):
        T, B = reward.shape
        dev = mu.device
        out = torch.empty(4, dtype=torch.float32, device=dev)
        lp = torch.empty(T, B, dtype=torch.float32, device=dev)
        cpg = torch.empty(T, B, dtype=torch.float32, device=dev)
        dv = torch.empty(T, B, dtype=torch.float32, device=dev)
        with on_device(dev):
            ws = workspace(dev)
            rc = lib().b200rl_vtrace_continuous_fwd(
                ptr(mu), ptr(sigma), ptr(mu_b), ptr(sigma_b), ptr(action), ptr(value), ptr(reward), ptr(weight), T, B, D,
                gamma, lambda_, rho_clip, c_clip, rho_pg_clip, ptr(out), ptr(lp), ptr(cpg), ptr(dv), ptr(ws),
                ws.numel() * 4, stream_ptr())
        _lib.check(rc, 'b200rl_vtrace_continuous_fwd')
        ctx.save_for_backward(mu, sigma, action, weight, cpg, dv)
        ctx.cfg = (T, B, D)
        ctx.set_materialize_grads(False)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_p, g_v, g_e):
        mu, sigma, action, weight, cpg, dv = ctx.saved_tensors
        T, B, D = ctx.cfg
        keep = [_g(x) for x in (g_p, g_v, g_e)]
        gm, gs = torch.empty_like(mu), torch.empty_like(sigma)
        gv = torch.empty(T + 1, B, dtype=torch.float32, device=mu.device)
        with on_device(mu.device):
            rc = lib().b200rl_vtrace_continuous_bwd(ptr(mu), ptr(sigma), ptr(action), ptr(weight), ptr(cpg), ptr(dv),
                                                    keep[0][1], keep[1][1], keep[2][1], T, B, D, ptr(gm), ptr(gs), ptr(gv),
                                                    stream_ptr())
        _lib.check(rc, 'b200rl_vtrace_continuous_bwd')
        return (gm, gs, gv) + (None, ) * 11
