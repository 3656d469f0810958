"""Pin the CPU oracle (oracle/gptq_oracle.py) against vectors produced by the reference itself
(tests/golden/*.npz, written by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import gptq_oracle as O
from conftest import same_cpu_as_golden

STRICT = same_cpu_as_golden()   # MKL picks kernels per CPU model: bit-exactness only promised on the generating CPU


@pytest.fixture(autouse=True)
def _one_thread():
    """The fixtures were generated with one MKL thread (PROVENANCE.json); sgemm rounding depends on it."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _qp(c):
    return O.QParams(bits=c["bits"], perchannel=c.get("perchannel", True), sym=c.get("sym", False),
                     mse=c.get("mse", False), trits=c.get("trits", False))


def _hessian(z, meta):
    st = O.HessianState(meta["K"])
    for x in torch.from_numpy(z["X"]):
        O.add_batch(st, x)
    return st


@pytest.mark.parametrize("tag", ["A", "B"])
def test_add_batch_matches_reference(tag, golden_A, golden_B):
    z, meta = golden_A if tag == "A" else golden_B
    st = _hessian(z, meta)
    Href = torch.from_numpy(z["H"])
    assert st.nsamples == meta["nsamples"]
    if STRICT:
        assert torch.equal(st.H, Href)
        assert torch.equal(st.last_input_mean, torch.from_numpy(z["input_mean"]))
    assert O.rel_fro(st.H, Href) < 1e-6
    # closed form (2/N) sum X^T X, the thing the CUDA syrk computes
    closed = O.hessian_closed_form([x for x in torch.from_numpy(z["X"])])
    assert O.rel_fro(Href, closed) < 1e-5


def _cases(tag):
    import json, os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"gptq_{tag}.npz"))
    return [(tag, c["name"]) for c in json.loads(bytes(z["meta"]).decode())["cases"]]


@pytest.mark.parametrize("tag,name", _cases("A") + _cases("B"))
def test_fasterquant_matches_reference(tag, name, golden_A, golden_B):
    z, meta = golden_A if tag == "A" else golden_B
    c = next(c for c in meta["cases"] if c["name"] == name)
    W = torch.from_numpy(z["W"])
    H = torch.from_numpy(z["H"])
    r = O.fasterquant(W, H, _qp(c), blocksize=128, percdamp=c.get("percdamp", .01),
                      groupsize=c.get("groupsize", -1), actorder=c.get("actorder", False),
                      static_groups=c.get("static_groups", False))
    Qref = torch.from_numpy(z[f"{name}/Q"])
    Q16 = r.Q.half()
    mism = float((Q16 != Qref).double().mean())
    if STRICT:
        assert mism == 0.0
        assert torch.equal(r.final_scale, torch.from_numpy(z[f"{name}/scale"]))
        assert torch.equal(r.final_zero, torch.from_numpy(z[f"{name}/zero"]))
    assert mism <= 1e-3
    assert abs(r.error - c["error"]) <= 1e-2 * abs(c["error"]) + 1e-12
    if not c.get("trits", False):
        # the integer codes the oracle reports dequantise to exactly the reference's Q
        g = r.g_idx
        s = r.scales[:, g]
        zr = r.zeros[:, g]
        deq = (s * (r.codes.float() - zr))
        assert torch.equal(deq, r.Q)


def test_pack3_matches_reference(golden_A):
    z, meta = golden_A
    W = torch.from_numpy(z["b3_perchannel/Q"])
    scale = torch.from_numpy(z["b3_perchannel/scale"])
    zero = torch.from_numpy(z["b3_perchannel/zero"])
    iw = O.intweight_from_linear(W, scale, zero)
    assert iw.max() <= 7
    qw = O.pack3(iw)
    assert np.array_equal(qw, z["b3_perchannel/qweight3"])
    assert np.array_equal(O.pack3_reference_loop(iw), z["b3_perchannel/qweight3"])
    assert np.array_equal(O.unpack(qw, 3, meta["K"]), iw)


def test_pack4_matches_reference(golden_A):
    z, meta = golden_A
    W = torch.from_numpy(z["b4_perchannel/Q"])
    iw = O.intweight_from_linear(W, torch.from_numpy(z["b4_perchannel/scale"]),
                                 torch.from_numpy(z["b4_perchannel/zero"]))
    qw = O.pack_pow2(iw, 4)
    assert np.array_equal(qw, z["b4_perchannel/qweight4"])
    assert np.array_equal(O.unpack(qw, 4, meta["K"]), iw)


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_pack_roundtrip_random(bits):
    rng = np.random.default_rng(bits)
    iw = rng.integers(0, 2 ** bits, size=(96, 40), dtype=np.uint32)
    assert np.array_equal(O.unpack(O.pack(iw, bits), bits, 96), iw)


def test_quantizer_matches_reference(golden_quantizer):
    z, meta = golden_quantizer
    for c in meta:
        x = torch.from_numpy(z[c["x"]]).clone()
        qp = _qp(c)
        s, zr = O.find_params(x, qp, weight=c.get("weight", True))
        assert torch.equal(s, torch.from_numpy(z[c["name"] + "/scale"])), c["name"]
        assert torch.equal(zr, torch.from_numpy(z[c["name"] + "/zero"])), c["name"]
        xq = O.quantize(torch.from_numpy(z[c["x"]]), s, zr, qp.maxq)
        assert torch.equal(xq, torch.from_numpy(z[c["name"] + "/xq"])), c["name"]
