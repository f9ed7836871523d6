"""Collector-side batch preparation: list of transition dicts -> the device tensors the learner's
operators take, through ONE pinned staging buffer and ONE host-to-device copy.

``preprocess_learn`` has the semantics of ``default_preprocess_learn`` (ding/policy/common_utils.py:28-98) on top of
``default_collate`` (ding/utils/data/collate_fn.py:80-160) for the transition dicts DI-engine's collectors produce (values:
tensors, numpy arrays, python numbers / bools, and dicts of those, e.g. a dict observation).  The reference stacks every field
into a fresh pageable tensor and moves the fields to the GPU one by one (``to_device``, ding/torch_utils/data_helper.py:22);
here every field is stacked STRAIGHT INTO its slice of one pinned buffer (``torch.stack(..., out=view)``: one pass over the
samples, no intermediate), the buffer crosses PCIe in one ``cudaMemcpyAsync``, and the per-field fix-ups (``done.float()``,
the n-step reward transpose, squeezes) run on device views.  Flags travel as one byte.  Staging buffers are cached per batch
layout, as the reference caches its plugin instances per shape (ding/hpc_rl/wrapper.py:74-83).  Host logic only.
"""
from collections import OrderedDict

import numpy as np
import torch

from .data import PackedBatch

_SLOTS = {}   # (device, layout signature) -> [PackedBatch, PackedBatch] (double-buffered: batch i+1 is staged while i is in use)
_TURN = {}
MAX_LAYOUTS = 8  # staging-buffer pairs kept alive (one pair per distinct batch layout, e.g. a smaller last minibatch)


def _leaf_spec(elem, cat_1dim):
    """shape / dtype of one collated field from its first sample, following default_collate's rules"""
    if isinstance(elem, torch.Tensor):
        shape = () if (tuple(elem.shape) == (1, ) and cat_1dim) else tuple(elem.shape)
        return shape, elem.dtype
    if isinstance(elem, np.ndarray):
        t = torch.as_tensor(elem)
        shape = () if (tuple(t.shape) == (1, ) and cat_1dim) else tuple(t.shape)
        return shape, t.dtype
    if isinstance(elem, (bool, np.bool_)):
        return (), torch.bool
    if isinstance(elem, (float, np.floating)):
        return (), torch.float32 if isinstance(elem, float) else torch.as_tensor(elem).dtype
    if isinstance(elem, (int, np.integer)):
        return (), torch.int64
    raise TypeError("collate: unsupported field type %s" % type(elem).__name__)


def _flatten(sample, prefix=()):
    for k, v in sample.items():
        if isinstance(k, str) and k.startswith('collate_ignore'):
            continue
        if isinstance(v, dict):
            yield from _flatten(v, prefix + (k, ))
        elif v is None:
            yield prefix + (k, ), None
        else:
            yield prefix + (k, ), v


def _get(sample, path):
    for k in path:
        sample = sample[k]
    return sample


def collate(data, device, cat_1dim=True, stream=None):
    """``default_collate`` of a list of (nested) transition dicts onto ``device``: {key: tensor with leading dim B} (nested dicts
    preserved), one pinned staging buffer, one H2D copy.  Returns (batch dict, event of the copy or None)."""
    B = len(data)
    first = data[0]
    paths, like = [], OrderedDict()
    for path, v in _flatten(first):
        paths.append(path)
        if v is None:
            like[path] = None
            continue
        shape, dtype = _leaf_spec(v, cat_1dim)
        like[path] = (tuple((B, ) + shape), dtype)
    device = torch.device(device)
    sig = (str(device), tuple((p, spec) for p, spec in like.items()))
    slots = _SLOTS.get(sig)
    if slots is None:
        tmpl = OrderedDict((p, None if spec is None else torch.zeros(spec[0], dtype=spec[1])) for p, spec in like.items())
        slots = [PackedBatch(tmpl, device), PackedBatch(tmpl, device)]
        if len(_SLOTS) >= MAX_LAYOUTS:  # pinned memory is a scarce resource: forget the least recently created layout
            old = next(iter(_SLOTS))
            del _SLOTS[old], _TURN[old]
        _SLOTS[sig] = slots
        _TURN[sig] = 0
    slot = slots[_TURN[sig]]
    _TURN[sig] ^= 1
    for path, spec in like.items():
        if spec is None:
            continue
        view = slot.host[path]
        leaf0 = _get(first, path)
        if isinstance(leaf0, torch.Tensor):
            vals = [_get(d, path) for d in data]
            if view.dim() == 1 and leaf0.dim() == 1:
                torch.cat(vals, 0, out=view)        # (1,) samples -> (B,)   (collate_fn.py:133-135)
            else:
                torch.stack(vals, 0, out=view)
        elif isinstance(leaf0, np.ndarray):
            vals = [torch.as_tensor(_get(d, path)) for d in data]
            if view.dim() == 1 and vals[0].dim() == 1:
                torch.cat(vals, 0, out=view)
            else:
                torch.stack(vals, 0, out=view)
        else:
            view.copy_(torch.as_tensor([_get(d, path) for d in data], dtype=view.dtype))
    flat, ev = slot.upload(stream)
    out = {}
    for path in paths:
        node = out
        for k in path[:-1]:
            node = node.setdefault(k, {})
        node[path[-1]] = flat[path]
    return out, ev


def preprocess_learn(data, device, use_priority_IS_weight=False, use_priority=False, use_nstep=False, ignore_done=False,
                     stream=None):
    """``default_preprocess_learn`` (ding/policy/common_utils.py:28-98) with the batch delivered on ``device``: collation as
    ``collate`` above, then the reference's fix-ups line for line on the device views.  The copy is enqueued on ``stream``
    (default: current) and the current stream waits for it."""
    elem = data[0]
    act = elem['action']
    discrete = isinstance(act, (np.ndarray, torch.Tensor)) and act.dtype in (np.int64, torch.int64)
    batch, ev = collate(data, device, cat_1dim=discrete, stream=stream)
    if ev is not None:
        torch.cuda.current_stream(torch.device(device)).wait_event(ev)
    for k in ('value', 'adv'):
        if k in batch and batch[k].dim() == 2 and batch[k].shape[1] == 1:
            batch[k] = batch[k].squeeze(-1)
    if ignore_done:
        batch['done'] = torch.zeros_like(batch['done']).float()
    else:
        batch['done'] = batch['done'].float()
    if batch['done'].dim() == 2 and batch['done'].shape[1] == 1:
        batch['done'] = batch['done'].squeeze(-1)
    if use_priority_IS_weight:
        assert use_priority, "Use IS Weight correction, but Priority is not used."
    if use_priority and use_priority_IS_weight:
        batch['weight'] = batch['priority_IS'] if 'priority_IS' in batch else batch['IS']
    else:
        batch['weight'] = batch.get('weight', None)
    if use_nstep:
        reward = batch['reward']
        if len(reward.shape) == 1:
            reward = reward.unsqueeze(1)
        if reward.ndim == 2:      # (batch_size, nstep) -> (nstep, batch_size)
            batch['reward'] = reward.transpose(0, 1).contiguous()
        elif reward.ndim == 3:    # (batch_size, agent_dim, nstep) -> (nstep, batch_size, agent_dim)
            batch['reward'] = reward.permute(2, 0, 1).contiguous()
        else:
            raise ValueError("The 'reward' tensor must be either 2D or 3D. Got shape: {}".format(reward.shape))
    else:
        if batch['reward'].dim() == 2 and batch['reward'].shape[1] == 1:
            batch['reward'] = batch['reward'].squeeze(-1)
    return batch


def get_gae(data, last_value, gamma, gae_lambda, device=None):
    """``Adder.get_gae`` (ding/rl_utils/adder.py:20-58): stacked values / rewards of one trajectory piece -> GAE on the device ->
    ``data[i]['adv']`` (host tensors, as the collector expects)."""
    from .rl_utils import gae, gae_data
    value = torch.stack([d['value'] for d in data])
    next_value = torch.stack([d['value'] for d in data][1:] + [last_value])
    reward = torch.stack([d['reward'] for d in data])
    if device is not None:
        value, next_value, reward = value.to(device), next_value.to(device), reward.to(device)
    adv = gae(gae_data(value, next_value, reward, None, None), gamma, gae_lambda)
    adv = adv.cpu()
    for i in range(len(data)):
        data[i]['adv'] = adv[i]
    return data


def nstep_return_data(reward, done, nstep, gamma=0.99, cum_reward=False, correct_terminate_gamma=True):
    """``Adder.get_nstep_return_data`` (ding/rl_utils/adder.py:97-155) for ONE trajectory held as stacked tensors instead of a
    deque of dicts: ``reward`` (T, ..., 1) -- the collector's per-step reward of shape (1,) or (agent_num, 1) stacked along T --
    and ``done`` (T,).  Returns ``(reward_n, done_n, value_gamma, next_index)``:

    * ``reward_n``  (T, ..., nstep): step i holds rewards i .. i+nstep-1, zero-padded past the end of the trajectory (the
      reference's ``fake_reward``); with ``cum_reward`` (T, ..., 1): ``sum_j gamma**j * reward[i+j]`` over the steps that exist;
    * ``done_n``    (T,): ``done[i+nstep-1]``, the last step's flag for the tail;
    * ``value_gamma`` (T,) fp32: ``gamma**nstep``, ``gamma**(T-i-1)`` for the tail (only meaningful with
      ``correct_terminate_gamma``, returned either way);
    * ``next_index`` (T,) int64: ``next_obs`` of step i is ``obs[next_index[i]]`` for i < T - nstep and the trajectory's last
      ``next_obs`` (index T, i.e. one past the stacked observations) for the tail.

    Index arithmetic only (gathers over a padded copy): works on host or device tensors alike."""
    T = reward.shape[0]
    if nstep == 1:
        ar = torch.arange(T, device=reward.device)
        return reward, done, torch.full((T, ), float(gamma), device=reward.device), ar + 1
    dev = reward.device
    idx = torch.arange(T, device=dev).unsqueeze(1) + torch.arange(nstep, device=dev).unsqueeze(0)  # (T, nstep): step i + j
    valid = idx < T
    pad = torch.cat([reward, torch.zeros((nstep, ) + tuple(reward.shape[1:]), dtype=reward.dtype, device=dev)], 0)
    win = pad[idx.clamp(max=T + nstep - 1)]                      # (T, nstep, ..., 1)
    win = win.movedim(1, -1).squeeze(-2)                         # (T, ..., nstep)
    if cum_reward:
        # the reference sums python-side: data[i]['reward'] * gamma**0 + data[i+1]['reward'] * gamma**1 + ... in that order
        out = torch.zeros_like(reward)
        for j in range(nstep):
            term = win[..., j:j + 1] * (gamma ** j)
            out = term if j == 0 else out + term
        reward_n = out
    else:
        reward_n = win
    last = torch.clamp(torch.arange(T, device=dev) + nstep - 1, max=T - 1)
    done_n = done[last]
    steps_left = (T - 1 - torch.arange(T, device=dev)).clamp(max=nstep).to(torch.float32)
    tail = torch.arange(T, device=dev) >= max(0, T - nstep)
    value_gamma = torch.where(tail, torch.pow(torch.tensor(float(gamma), device=dev), steps_left),
                              torch.full((T, ), float(gamma) ** nstep, device=dev))
    next_index = torch.where(tail, torch.full((T, ), T, device=dev), torch.arange(T, device=dev) + nstep)
    return reward_n, done_n, value_gamma, next_index


def _transpose_steps(steps):
    """list of per-step dicts -> dict of per-key lists (one level of nested dicts transposed too, except ``prev_state``): what
    ``lists_to_dicts(..., recursive=True)`` does for the sequence samples (ding/utils/default_helper.py:41-76)"""
    if len(steps) == 0:
        raise ValueError("empty data")
    out = {}
    for k, v0 in steps[0].items():
        col = [s[k] for s in steps]
        if isinstance(v0, dict) and k != 'prev_state':
            out[k] = {kk: [c[kk] for c in col] for kk in v0.keys()}
        else:
            out[k] = col
    return out


def get_train_sample(data, unroll_len, last_fn_type='last', null_transition=None):
    """``Adder.get_train_sample`` (ding/rl_utils/adder.py:158-232): cut one trajectory (list of transition dicts) into sequence
    samples of ``unroll_len`` steps, each returned as a dict of per-key lists.  The trailing remainder is handled as the reference
    does: ``'drop'`` discards it; ``'last'`` completes it with the last steps of the previous piece put in FRONT of it (or, when
    there is no previous piece, pads BEHIND it with null transitions); ``'null_padding'`` always pads behind.  A null transition
    is a deep copy of ``null_transition`` if given, else of the remainder's first step with ``null=True``, zeroed obs / action /
    reward, ``done=True`` and ``value_gamma=0``.  Host-side list surgery, as in the reference (no tensors are computed)."""
    import copy
    if unroll_len == 1:
        return data
    n_full = len(data) // unroll_len
    pieces = [data[i * unroll_len:(i + 1) * unroll_len] for i in range(n_full)]
    rest = data[n_full * unroll_len:] if n_full * unroll_len < len(data) else None
    if rest is not None:
        missing = unroll_len - len(rest)

        def nulls():
            tmpl = copy.deepcopy(rest[0])
            tmpl['null'] = True
            obs = tmpl['obs']
            tmpl['obs'] = {k: torch.zeros_like(v) for k, v in obs.items()} if isinstance(obs, dict) else torch.zeros_like(obs)
            if 'action' in tmpl:
                tmpl['action'] = torch.zeros_like(tmpl['action'])
            tmpl['done'] = True
            tmpl['reward'] = torch.zeros_like(tmpl['reward'])
            if 'value_gamma' in tmpl:
                tmpl['value_gamma'] = 0.
            src = null_transition if null_transition is not None else tmpl
            return [copy.deepcopy(src) for _ in range(missing)]

        if last_fn_type == 'last' and pieces:
            pieces.append(copy.deepcopy(pieces[-1][-missing:]) + rest)
        elif last_fn_type in ('last', 'null_padding'):
            pieces.append(rest + nulls())
        elif last_fn_type != 'drop':
            raise ValueError("last_fn_type should be in ['last', 'drop', 'null_padding'], got %r" % (last_fn_type, ))
    return [_transpose_steps(p) for p in pieces]
