"""Host-side staging of a learner batch: ONE pinned host buffer and ONE device buffer per batch slot.

The operators take separate tensors, as ``ding.rl_utils`` does; what crosses PCIe per step is a dozen of them
(``default_collate`` + ``to_device``, ding/policy/common_utils.py:28-98, ding/torch_utils/data_helper.py:543).  ``PackedBatch``
lays those tensors out back to back (256-byte aligned) in one pinned allocation, so the collector writes into views of it
and the H2D transfer of a step is a single ``cudaMemcpyAsync`` instead of one per tensor; the device side hands out views
of one allocation with the original shapes and dtypes.  Plumbing only -- no arithmetic.

``narrow``: the learner step is PCIe-bound (46 MB per 16 us of kernels), and a third of a transition's index / flag bytes are
padding: a discrete action fits one byte, ``done`` / ``traj_flag`` are 0 / 1.  ``narrow={'action': torch.uint8, ...}`` keeps
those fields in the narrow dtype ON THE WIRE (host views and the device staging buffer) and widens them to the dtype the
operators take (int64 / float32, exact for these values) on the device after the copy: 88 -> 75 bytes per transition at
config D.  The caller is responsible for the values fitting the narrow dtype (checked when the template is copied in).
"""
from collections import OrderedDict

import torch

_ALIGN = 256


class PackedBatch:

    def __init__(self, like: dict, device, narrow=None):
        """``like``: name -> tensor (shape / dtype template, contents are copied in); None entries are kept as None.
        ``narrow``: name -> wire dtype for fields that travel narrower than the operators take them (see above)."""
        self.device = torch.device(device)
        self.layout = OrderedDict()
        self.widen = {}
        off = 0
        for k, v in like.items():
            if v is None:
                self.layout[k] = None
                continue
            wire = (narrow or {}).get(k, v.dtype)
            if wire != v.dtype:
                if not torch.equal(v.to(wire).to(v.dtype), v):
                    raise ValueError("PackedBatch: field %r does not survive the wire dtype %s" % (k, wire))
                self.widen[k] = v.dtype
            nbytes = v.numel() * torch.empty((), dtype=wire).element_size()
            self.layout[k] = (off, nbytes, wire, tuple(v.shape))
            off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = off
        self.host_buf = torch.empty(max(off, 1), dtype=torch.uint8)
        if self.device.type == 'cuda':
            self.host_buf = self.host_buf.pin_memory()
        self.dev_buf = torch.empty(max(off, 1), dtype=torch.uint8, device=self.device)
        self.host = self._views(self.host_buf)
        for k, v in like.items():
            if v is not None:
                self.host[k].copy_(v)

    def _views(self, buf):
        out = OrderedDict()
        for k, spec in self.layout.items():
            if spec is None:
                out[k] = None
            else:
                off, nbytes, dtype, shape = spec
                out[k] = buf[off:off + nbytes].view(dtype).view(shape)
        return out

    def payload_bytes(self):
        return sum(spec[1] for spec in self.layout.values() if spec is not None)

    def upload(self, stream=None):
        """Enqueue the single H2D copy on ``stream`` (default: the current one); returns (fresh device views, event)."""
        if self.device.type != 'cuda':  # layout tests on a host without a GPU: plain copy, nothing to wait for
            self.dev_buf.copy_(self.host_buf)
            views = self._views(self.dev_buf)
            for k, dt in self.widen.items():
                views[k] = views[k].to(dt)
            return views, None
        stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            self.dev_buf.copy_(self.host_buf, non_blocking=True)
            views = self._views(self.dev_buf)
            for k, dt in self.widen.items():  # dtype widening on the device, on the copy stream (exact: 0/1 flags, small ints)
                views[k] = views[k].to(dt)
            ev = torch.cuda.Event()
            ev.record(stream)
        return views, ev
