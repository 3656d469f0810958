"""Place the UNMODIFIED reference (JerryZhou54/gptq) modules of the hot path under baseline/_ref/.

    python baseline/install_reference.py            # run in the build container (needs /root/reference, read-only)

The reference is a flat set of scripts without packaging metadata (its only setup file, setup_cuda.py, builds the
quant_cuda extension), so "installing" it means copying, byte for byte, the modules `import gptq` needs:
gptq.py, quant.py, modelutils.py, plot_activation.py and the bcq_quant package (gptq.py:10-14).  baseline/_ref/ is
git-ignored (the reference's sources never enter this repository's history) but travels to the GPU box with the snapshot,
where /root/reference does not exist.  baseline/reference.py imports it with the two shims of SURVEY 8c and never edits it.
A manifest with the sha256 of every file is written beside the copies.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("GPTQ_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["gptq.py", "quant.py", "modelutils.py", "plot_activation.py", "LICENSE"]
DIRS = ["bcq_quant"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(force: bool = False) -> str:
    """Copy the modules; returns the destination.  No-op (keeps what is there) when the reference tree is absent."""
    if not os.path.isdir(REF):
        if os.path.exists(os.path.join(DST, "gptq.py")):
            return DST
        raise FileNotFoundError(f"{REF} not found and {DST} is empty: run this script in the build container")
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(REF, name)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(DST, name))
            manifest[name] = _sha(src)
    for d in DIRS:
        src = os.path.join(REF, d)
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for root, _, files in os.walk(dst):
            for f in sorted(files):
                p = os.path.join(root, f)
                manifest[os.path.relpath(p, DST)] = _sha(p)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
