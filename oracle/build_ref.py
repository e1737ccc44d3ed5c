"""TEST / BASELINE INFRASTRUCTURE ONLY -- build recipe for ``oracle/_ref``.

The reference is pure Python; its "build" is byte-compilation.  This recipe compiles the reference package from the
sources WHERE THEY LIE under ``/root/reference`` (nothing is copied) and writes only the outputs -- byte-compiled
modules, stored as ``module.refc`` (a ``.pyc`` under another suffix: the gpurun snapshot drops ``*.pyc``) -- into
``oracle/_ref/DistantSpeech/...``; ``oracle/ref_harness.py`` installs a small import hook that loads them.  ``oracle/_ref/`` is git-ignored (it never enters the history) but
not gpurun-ignored, so the compiled reference travels to the GPU box like the repo's own ``.so`` and can be

  * timed there as the CPU baseline of ``bench.py`` (``cpu_baseline.kind = "reference"``), and
  * imported by ``tests/test_oracle_vs_reference.py`` to re-pin the NumPy oracle on that box too.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present (the build container); the GPU box only
uses the prebuilt files.  Import goes through ``oracle/ref_harness.py`` (``DS_REFERENCE_ROOT=oracle/_ref``), which
applies the same stubs / compatibility patches as for the source tree.

    python oracle/build_ref.py
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("DS_REFERENCE_SRC", "/root/reference")
OUT_ROOT = os.path.join(HERE, "_ref")
PACKAGE = "DistantSpeech"


def available() -> bool:
    """True when a compiled reference is present (marker written by build())."""
    return os.path.exists(os.path.join(OUT_ROOT, PACKAGE, "transform", "transform.refc"))


def build(verbose: bool = False) -> int:
    """Byte-compile every module of the reference package into oracle/_ref; returns the number of modules."""
    src_pkg = os.path.join(SRC_ROOT, PACKAGE)
    if not os.path.isdir(src_pkg):
        raise RuntimeError("reference sources not present at %s" % src_pkg)
    import shutil
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)          # the reference's own docstring escapes
    shutil.rmtree(os.path.join(OUT_ROOT, PACKAGE), ignore_errors=True)
    n = 0
    for dirpath, dirnames, filenames in os.walk(src_pkg):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC_ROOT)
        for f in filenames:
            if not f.endswith(".py"):
                continue
            src = os.path.join(dirpath, f)
            dst = os.path.join(OUT_ROOT, rel, f[:-3] + ".refc")        # byte code next to where module.py would be
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            try:
                # dfile: the path recorded in tracebacks / co_filename -- the reference's own location, for citations
                py_compile.compile(src, cfile=dst, dfile=os.path.join("/root/reference", rel, f), doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
                n += 1
            except py_compile.PyCompileError as e:                      # a module that does not parse on this Python
                if verbose:
                    sys.stderr.write("skipped %s: %s\n" % (src, e.msg.strip().splitlines()[-1]))
    with open(os.path.join(OUT_ROOT, "BUILD_INFO.txt"), "w") as fh:
        fh.write("byte-compiled from %s by oracle/build_ref.py with Python %s; %d modules; no sources copied\n"
                 % (src_pkg, sys.version.split()[0], n))
    return n


if __name__ == "__main__":
    print("compiled %d reference modules into %s" % (build(verbose=True), OUT_ROOT))
