"""TEST / BENCH INFRASTRUCTURE ONLY -- builds ``oracle/_ref/ding_hotpath.zip`` from the UNMODIFIED reference.

The reference's hot path is pure Python, so "compiling the reference from the sources where they lie" (the rule for a
C reference: outputs only into ``oracle/_ref/``, no reference sources in the repository) means byte-compiling it: every
file below is read from ``/root/reference`` and written as a sourceless ``.pyc`` into ONE zip archive that python
imports through ``zipimport``.  ``oracle/_ref/`` is git-ignored (stays out of history) but not gpurun-ignored, so the
archive travels to the GPU box like the built ``.so`` -- there ``bench.py --impl reference`` and the ``cpu_baseline``
leg time the reference's own functions (``cpu_baseline.kind = "reference"``) instead of the oracle port, and
``tests/test_reference_suite.py`` can check the product against the live reference next to the GPU.

    python oracle/make_ref.py            # (re)build, prints the archive path; also run by __graft_entry__.build()

The archive holds byte code of exactly these reference files (nothing of ours is mixed in):
"""
import io
import os
import py_compile
import sys
import tempfile
import warnings
import zipfile

REF_ROOT = os.environ.get("DI_ENGINE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "ding_hotpath.zip")

FILES = [
    "ding/__init__.py",                # enable_hpc_rl switch (ding/__init__.py:10)
    "ding/hpc_rl/__init__.py",
    "ding/hpc_rl/wrapper.py",          # the plugin boundary (hpc_wrapper :86-133)
    "ding/rl_utils/value_rescale.py",
    "ding/rl_utils/gae.py",
    "ding/rl_utils/td.py",
    "ding/rl_utils/ppo.py",
    "ding/rl_utils/isw.py",
    "ding/rl_utils/vtrace.py",
    "ding/rl_utils/upgo.py",
    "ding/rl_utils/a2c.py",            # sibling head
    "ding/rl_utils/retrace.py",        # ACER's return operator
    "ding/rl_utils/happo.py",          # HAPPO heads (ppo_error with the per-sample factor)
    "ding/rl_utils/acer.py",           # ACER heads
    "ding/rl_utils/ppg.py",            # PPG joint (auxiliary phase) loss
]


def available():
    return os.path.isfile(os.path.join(REF_ROOT, FILES[0]))


def stale():
    if not os.path.isfile(ARCHIVE):
        return True
    t = os.path.getmtime(ARCHIVE)
    return any(os.path.getmtime(os.path.join(REF_ROOT, f)) > t for f in FILES) or os.path.getmtime(__file__) > t


def build(force=False):
    """Byte-compile the reference files into the archive; returns its path (None when the reference tree is absent)."""
    if not available():
        return ARCHIVE if os.path.isfile(ARCHIVE) else None
    if not force and not stale():
        return ARCHIVE
    os.makedirs(OUT_DIR, exist_ok=True)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as z, tempfile.TemporaryDirectory() as tmp:
        for rel in FILES:
            src = os.path.join(REF_ROOT, rel)
            cfile = os.path.join(tmp, "m.pyc")
            # unchecked hash-based pyc: valid without the source file next to it
            with warnings.catch_warnings():  # the reference's docstrings hold a few invalid escape sequences
                warnings.simplefilter("ignore", SyntaxWarning)
                py_compile.compile(src, cfile=cfile, dfile=rel, doraise=True, optimize=0,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(cfile, rel[:-3] + ".pyc")
        z.writestr("MANIFEST.txt", "byte code (python %d.%d) of the unmodified reference files:\n%s\n" %
                   (sys.version_info[0], sys.version_info[1], "\n".join(FILES)))
    with open(ARCHIVE, "wb") as f:
        f.write(buf.getvalue())
    return ARCHIVE


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "reference tree not present at %s and no archive built earlier" % REF_ROOT)
