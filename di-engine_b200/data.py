"""Host-side staging of a learner batch: ONE pinned host buffer and ONE device buffer per batch slot.

The operators take separate tensors, as ``ding.rl_utils`` does; what crosses PCIe per step is a dozen of them
(``default_collate`` + ``to_device``, ding/policy/common_utils.py:28-98, ding/torch_utils/data_helper.py:543).  ``PackedBatch``
lays those tensors out back to back (256-byte aligned) in one pinned allocation, so the collector writes into views of it
and the H2D transfer of a step is a single ``cudaMemcpyAsync`` instead of one per tensor; the device side hands out views
of one allocation with the original shapes and dtypes.  Plumbing only -- no arithmetic.
"""
from collections import OrderedDict

import torch

_ALIGN = 256


class PackedBatch:

    def __init__(self, like: dict, device):
        """``like``: name -> tensor (shape / dtype template, contents are copied in); None entries are kept as None."""
        self.device = torch.device(device)
        self.layout = OrderedDict()
        off = 0
        for k, v in like.items():
            if v is None:
                self.layout[k] = None
                continue
            nbytes = v.numel() * v.element_size()
            self.layout[k] = (off, nbytes, v.dtype, tuple(v.shape))
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
            return self._views(self.dev_buf), None
        stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            self.dev_buf.copy_(self.host_buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        return self._views(self.dev_buf), ev
