"""Mint the golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py          # writes the fixtures that do not exist yet (new cases)
    python tests/golden/make_golden.py --all    # re-mints every fixture

For every case of ``tests/cases.build_cases()`` this runs the real ``ding.rl_utils`` functions (loaded read-only
from /root/reference by ``oracle/ref_loader.py``) on CPU fp32 and stores inputs, scalar parameters, forward
outputs and input-gradients in ``tests/golden/<case>.npz``.  The reference holds no golden vectors of its own for
this path so these files are the pins that travel to the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import cases  # noqa: E402


def _encode_params(params):
    out = {}
    for k, v in params.items():
        if isinstance(v, list):  # NGU list-gamma: list of 0-dim tensors
            out[k] = {'__tensor_list__': [float(x) for x in v]}
        else:
            out[k] = v
    return out


def main():
    torch.set_num_threads(1)
    ref = ref_loader.load()
    n = 0
    for name, (op, tensors, params) in cases.build_cases().items():
        if '--all' not in sys.argv and os.path.isfile(os.path.join(HERE, name + '.npz')):
            continue
        res = cases.run_api(ref, op, tensors, params)
        blob = {}
        meta = {'op': op, 'params': _encode_params(params), 'none_inputs': [], 'bool_inputs': [],
                'torch': torch.__version__}
        for k, v in tensors.items():
            if v is None:
                meta['none_inputs'].append(k)
            else:
                if v.dtype == torch.bool:
                    meta['bool_inputs'].append(k)
                blob['in_' + k] = v.numpy()
        blob.update(res)
        blob['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
        n += 1
    print('wrote %d fixtures to %s' % (n, HERE))


if __name__ == '__main__':
    main()
