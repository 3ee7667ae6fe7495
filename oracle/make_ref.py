#!/usr/bin/env python3
"""Recipe for oracle/_ref/: the UNMODIFIED reference implementation of the hot path, for use as the timed CPU arm
(`bench.py --impl reference`, cpu_baseline.kind = "reference") and as a second checker beside the numpy restatement.

The reference (eepperly/Randomly-Pivoted-Cholesky) is pure Python without packaging metadata, so "building" it is
copying the four modules of the path -- rpcholesky.py, matrix.py, kernels.py, lra.py -- byte for
byte from /root/reference into oracle/_ref/.  oracle/_ref/ is git-ignored (reference sources never enter the history)
but not gpurun-ignored, so the copy travels to the GPU box like the built .so files.  /root/reference exists only in the
build container: on the GPU box this script is a no-op and bench.py uses whatever copy travelled (or falls back to the
numpy port, kind = "port").

    python oracle/make_ref.py            # -> oracle/_ref/{rpcholesky,matrix,kernels,lra}.py + MANIFEST.json (sha256)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RPC_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
MODULES = ("rpcholesky.py", "matrix.py", "kernels.py", "lra.py")


def make(verbose=True):
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print("oracle/make_ref.py: %s not present (GPU box): keeping %s as it is" % (REF_ROOT, DEST))
        return os.path.isdir(DEST)
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in MODULES:
        src = os.path.join(REF_ROOT, name)
        shutil.copyfile(src, os.path.join(DEST, name))
        manifest[name] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest, "modified": False}, f, indent=1)
    if verbose:
        print("oracle/_ref: copied %s from %s" % (", ".join(MODULES), REF_ROOT))
    return True


def load():
    """Import the reference modules from oracle/_ref (None if the copy is absent).  They import each other by bare name
    (`from matrix import ...`), so the directory goes on sys.path; the names are removed again so that they cannot shadow
    anything else."""
    if not all(os.path.exists(os.path.join(DEST, m)) for m in MODULES):
        return None
    import importlib
    saved = {k: sys.modules.pop(k) for k in ("rpcholesky", "matrix", "kernels", "lra") if k in sys.modules}
    sys.path.insert(0, DEST)
    try:
        mods = {k: importlib.import_module(k) for k in ("kernels", "lra", "matrix", "rpcholesky")}
    finally:
        sys.path.remove(DEST)
        for k in ("rpcholesky", "matrix", "kernels", "lra"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return mods


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
