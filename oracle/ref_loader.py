"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* DI-engine hot-path files from /root/reference.

This exists only in the build container (the GPU box has no /root/reference). It is used to
  (1) validate the CPU restatement in ``oracle/rl_oracle.py`` and
  (2) mint the golden fixtures under ``tests/golden/`` (see ``tests/golden/make_golden.py``).
Nothing in the product package imports this module.

The hot-path files (ding/rl_utils/{gae,td,ppo,vtrace,upgo,isw,value_rescale}.py) only need torch, numpy,
``ding.hpc_rl`` (which needs ``ditk.logging``, ding/hpc_rl/wrapper.py:2) and an unused
``from ding.torch_utils import to_tensor`` (ding/rl_utils/td.py:12). We stub those three modules and register an
empty ``ding.rl_utils`` package whose ``__path__`` points at the reference directory so that the heavy package
``__init__`` (which pulls gym/easydict/treetensor...) is skipped.
"""
import importlib
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("DI_ENGINE_REFERENCE", "/root/reference")
_HOT_MODULES = ["value_rescale", "gae", "td", "ppo", "isw", "vtrace", "upgo"]
_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ding", "rl_utils", "gae.py"))


def load():
    """Return a namespace module holding every public name of the reference hot-path files."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "ditk" not in sys.modules:
        ditk = types.ModuleType("ditk")
        ditk.logging = logging
        sys.modules["ditk"] = ditk
        sys.modules["ditk.logging"] = logging
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import ding  # noqa: F401  (top-level __init__ needs only os + torch)
    if "ding.torch_utils" not in sys.modules:
        tu = types.ModuleType("ding.torch_utils")
        tu.to_tensor = lambda x, *a, **k: x
        sys.modules["ding.torch_utils"] = tu
    if "ding.rl_utils" not in sys.modules:
        pkg = types.ModuleType("ding.rl_utils")
        pkg.__path__ = [os.path.join(REF_ROOT, "ding", "rl_utils")]
        sys.modules["ding.rl_utils"] = pkg
    pkg = sys.modules["ding.rl_utils"]
    import ding.hpc_rl  # noqa: F401
    for m in _HOT_MODULES:
        mod = importlib.import_module("ding.rl_utils." + m)
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(pkg, k, v)
    _loaded = pkg
    return pkg
