"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/gptq_b200.h declares, and the Python mirror refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "gptq_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gptq_b200_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from gptq_b200 import _lib
    return _lib


def test_header_declares_expected_surface():
    syms = _header_symbols()
    for s in ("gptq_b200_syrk_accum", "gptq_b200_hinv_upper", "gptq_b200_colloop", "gptq_b200_pack_codes",
              "gptq_b200_find_params", "gptq_b200_quantize"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    dll = ctypes.CDLL(lib.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(dll, s), f"{s} declared in include/gptq_b200.h but not exported"
    assert set(lib.SIGNATURES) == set(_header_symbols())


def test_abi_version_and_error_string(lib):
    L = lib.load()
    assert L.gptq_b200_abi_version() == lib.ABI_VERSION == 3
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = L.gptq_b200_pack_codes(0, 4, 32, 32, 3, 0, 0)
    assert rc == -1
    assert b"argument 1" in L.gptq_b200_last_error()
    rc = L.gptq_b200_pack_codes(1, 4, 33, 33, 3, 1, 0)
    assert rc == -3
    rc = L.gptq_b200_hinv_upper(1, 100, 100, 2, 100, 1, 1, 0, 0)
    assert rc == -2
    assert L.gptq_b200_hinv_upper_workspace_bytes(256) == 128 * 128 * 4
    assert L.gptq_b200_colloop_workspace_bytes(1000, 4096, 128) == 1000 * 512 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    import torch.nn as nn
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer, quantize
    with pytest.raises(lib.GptqB200Error):
        GPTQ(nn.Linear(32, 8))
    q = Quantizer()
    q.configure(4, perchannel=True, sym=False)
    with pytest.raises(lib.GptqB200Error):
        q.find_params(torch.randn(8, 32))
    with pytest.raises(lib.GptqB200Error):
        quantize(torch.randn(8, 32), torch.ones(8, 1), torch.zeros(8, 1), 15)


def test_reference_api_surface():
    """Names and signatures the reference drivers rely on."""
    import inspect
    import gptq
    import quant
    sig = inspect.signature(gptq.GPTQ.fasterquant)
    for k, d in dict(blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False, model_name="opt",
                     layer_name="layer", lut_quant=False, non_linear_quant=False, columnwise=False).items():
        assert sig.parameters[k].default == d
    assert list(inspect.signature(gptq.GPTQ.add_batch).parameters) == ["self", "inp", "out"]
    csig = inspect.signature(quant.Quantizer.configure)
    assert [p for p in csig.parameters] == ["self", "bits", "perchannel", "sym", "mse", "norm", "grid", "maxshrink", "trits"]
    assert inspect.signature(quant.Quantizer.find_params).parameters["weight"].default is True
    assert list(inspect.signature(quant.Quant3Linear.pack).parameters)[:4] == ["self", "linear", "scales", "zeros"]
    q = quant.Quantizer()
    assert set(dict(q.named_buffers())) == {"maxq", "scale", "zero"}
    ql = quant.Quant3Linear(64, 16)
    assert ql.qweight.shape == (64 // 32 * 3, 16) and ql.qweight.dtype == torch.int32
    assert callable(quant.make_quant3) and callable(quant.quantize)
