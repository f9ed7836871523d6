"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* DI-engine hot-path files.

Source, in this order: the reference tree at /root/reference (build container only), else the byte-compiled archive
``oracle/_ref/ding_hotpath.zip`` that ``oracle/make_ref.py`` builds from that tree (git-ignored build output that
travels to the GPU box, where /root/reference does not exist).  It is used to
  (1) validate the CPU restatement in ``oracle/rl_oracle.py``,
  (2) mint the golden fixtures under ``tests/golden/`` (see ``tests/golden/make_golden.py``),
  (3) time the reference itself as ``bench.py``'s CPU arm (``cpu_baseline.kind = "reference"``).
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
import warnings

REF_ROOT = os.environ.get("DI_ENGINE_REFERENCE", "/root/reference")
ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ding_hotpath.zip")
_HOT_MODULES = ["value_rescale", "gae", "td", "ppo", "isw", "vtrace", "upgo", "a2c", "retrace", "happo", "acer", "ppg"]
_loaded = None


def tree_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ding", "rl_utils", "gae.py"))


def source():
    """'tree' (reference sources present), 'archive' (byte-compiled by oracle/make_ref.py) or None."""
    if os.environ.get("DI_ENGINE_REF_FORCE_ARCHIVE") != "1" and tree_available():
        return "tree"
    if os.path.isfile(ARCHIVE):
        return "archive"
    return None


def available() -> bool:
    return source() is not None


def load():
    """Return a namespace module holding every public name of the reference hot-path files."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = source()
    if src is None:
        raise RuntimeError("reference not available: no tree at %s and no archive %s" % (REF_ROOT, ARCHIVE))
    root = REF_ROOT if src == "tree" else ARCHIVE
    if "ditk" not in sys.modules:
        ditk = types.ModuleType("ditk")
        ditk.logging = logging
        sys.modules["ditk"] = ditk
        sys.modules["ditk.logging"] = logging
    if root not in sys.path:
        sys.path.insert(0, root)
    import ding  # noqa: F401  (top-level __init__ needs only os + torch)
    if "ding.torch_utils" not in sys.modules:
        tu = types.ModuleType("ding.torch_utils")
        tu.to_tensor = lambda x, *a, **k: x
        sys.modules["ding.torch_utils"] = tu
    if "ding.rl_utils" not in sys.modules:
        pkg = types.ModuleType("ding.rl_utils")
        pkg.__path__ = [os.path.join(root, "ding", "rl_utils")]
        sys.modules["ding.rl_utils"] = pkg
    pkg = sys.modules["ding.rl_utils"]
    import ding.hpc_rl  # noqa: F401
    for m in _HOT_MODULES:
        with warnings.catch_warnings():  # the reference's docstrings hold a few invalid escape sequences
            warnings.simplefilter("ignore", SyntaxWarning)
            mod = importlib.import_module("ding.rl_utils." + m)
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(pkg, k, v)
    _loaded = pkg
    return pkg
