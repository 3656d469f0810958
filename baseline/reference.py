"""Import the unmodified reference from baseline/_ref/ (see install_reference.py) -- BENCH / TEST INFRASTRUCTURE ONLY.

Two shims, neither touching the reference's files:
  1. gptq.py:14 imports plot_activation -> matplotlib, which is not installed: stub modules are injected;
  2. gptq.py:292 calls torch.cuda.synchronize() unconditionally: replaced by a no-op ONLY when no CUDA device exists.
The reference's module names (`gptq`, `quant`) collide with this repository's drop-in shims of the same names, so they are
imported from baseline/_ref with the path put first and then removed from sys.modules again; callers get module objects.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_NAMES = ("gptq", "quant", "plot_activation", "modelutils", "bcq_quant")
_cache = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "gptq.py"))


def load():
    """Returns (reference gptq module, reference quant module)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"{REF_DIR}/gptq.py missing: run `python baseline/install_reference.py` in the build container")
    import torch
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in _NAMES or k.startswith("bcq_quant.")}
    sys.path.insert(0, REF_DIR)
    try:
        # the reference prints "CUDA extension not installed." at import (quant.py, its optional quant_cuda module): keep the
        # bench's stdout to its one JSON line
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            import gptq as ref_gptq      # noqa: the reference's gptq.py
            import quant as ref_quant    # noqa
    finally:
        sys.path.remove(REF_DIR)
        for k in list(sys.modules):
            if k in _NAMES or k.startswith("bcq_quant."):
                del sys.modules[k]
        sys.modules.update(saved)
    assert os.path.dirname(os.path.abspath(ref_gptq.__file__)) == REF_DIR
    _cache = (ref_gptq, ref_quant)
    return _cache
