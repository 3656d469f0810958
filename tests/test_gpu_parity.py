"""GPU parity tests: every C-ABI entry point of libgptq_b200.so against the CPU oracle and the
reference-generated golden vectors.  Run on the B200 box: ``pytest -m gpu``.

Tolerances (BASELINE.json north_star): packed words bit-exact; H within 1e-5 relative Frobenius;
codes / dequantised weights agree on >= 99.9 % of entries; sum(Losses) within 1 %.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import gptq_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gptq_b200 import _lib
    from gptq_b200.gptq import GPTQ
    from gptq_b200 import quant as Q
    DEV = torch.device("cuda:0")


def _log(**kw):
    print("  [metric] " + json.dumps(kw))


# |frac(w/scale) - 0.5| below which a rounding flip between the CUDA path and the oracle counts as a tie (see
# tests/test_gpu_headline.py): same_H = both solvers see the oracle's Hessian, gpu_H = H from the tcgen05 syrk
TIE_TOL = {"same_H": 2.0 ** -18, "gpu_H": 2.0 ** -14}


# ------------------------------------------------------------------------------------------ pack
@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("R,K", [(40, 96), (64, 128), (200, 352), (1024, 1024)])
def test_pack_codes_bit_exact(bits, R, K):
    rng = np.random.default_rng(R * 131 + K + bits)
    codes = rng.integers(0, 2 ** bits, size=(R, K), dtype=np.uint8)
    want = O.pack(np.ascontiguousarray(codes.T).astype(np.uint32), bits)
    got = Q.pack_codes(torch.from_numpy(codes).to(DEV), bits)
    assert np.array_equal(got.cpu().numpy(), want)
    back = Q.unpack_codes(got, bits)
    assert np.array_equal(back.cpu().numpy(), codes)


def test_pack_roundtrip_full_size():
    # Llama-2-7B down_proj shape; size-independent property: unpack(pack(c)) == c, and a checksum of the words
    R, K = 4096, 11008
    g = torch.Generator(device="cuda").manual_seed(0)
    for bits in (2, 3, 4):
        codes = torch.randint(0, 2 ** bits, (R, K), device=DEV, dtype=torch.uint8, generator=g)
        qw = Q.pack_codes(codes, bits)
        assert qw.shape == (K * bits // 32, R)
        assert torch.equal(Q.unpack_codes(qw, bits), codes)
        # popcount(words) == sum over codes of popcount(code)
        pc_words = sum(int(((qw.view(torch.int32) >> s) & 1).sum()) for s in range(32))
        pc_codes = sum(int(((codes >> s) & 1).sum()) for s in range(bits))
        assert pc_words == pc_codes


def test_pack_weights_matches_reference(golden_A):
    z, meta = golden_A
    for name, bits, key in (("b3_perchannel", 3, "qweight3"), ("b4_perchannel", 4, "qweight4")):
        lin = nn.Linear(meta["K"], meta["R"], bias=False)
        lin.weight.data = torch.from_numpy(z[f"{name}/Q"])             # CPU fp16, as the reference driver hands it over
        ql = Q.QuantLinear(bits, meta["K"], meta["R"])
        ql.pack(lin, torch.from_numpy(z[f"{name}/scale"]), torch.from_numpy(z[f"{name}/zero"]))
        assert ql.qweight.device.type == "cpu"
        assert np.array_equal(ql.qweight.numpy(), z[f"{name}/{key}"])
        if bits == 3:
            assert np.array_equal(ql.zeros.numpy(), z[f"{name}/pack_zeros"])


# ------------------------------------------------------------------------------------------ quantizer
def test_quantizer_matches_reference(golden_quantizer):
    z, meta = golden_quantizer
    for c in meta:
        q = Q.Quantizer()
        q.configure(c["bits"], perchannel=c["perchannel"], sym=c["sym"], mse=c.get("mse", False), trits=c.get("trits", False))
        x = torch.from_numpy(z[c["x"]]).to(DEV)
        q.find_params(x.clone(), weight=c.get("weight", True))
        s_ref = torch.from_numpy(z[c["name"] + "/scale"])
        z_ref = torch.from_numpy(z[c["name"] + "/zero"])
        assert q.scale.shape == s_ref.shape, c["name"]
        if c.get("mse", False):
            # |q-x|^2.4 uses CUDA powf vs torch CPU pow: the argmin over the 80 shrink steps may move between two steps
            # whose errors are equal to rounding.  Every row whose scale differs must be such a near-tie: the error of the
            # GPU's choice, recomputed with the oracle's arithmetic, is within 1e-5 of the error of the oracle's choice.
            maxq = 2 ** c["bits"] - 1
            xr = x.cpu().float().flatten(1) if c["perchannel"] else x.cpu().float().flatten().unsqueeze(0)
            def _err(sc, zr):
                sc, zr = sc.cpu().reshape(-1, 1)[: xr.shape[0]], zr.cpu().reshape(-1, 1)[: xr.shape[0]]
                qx = O.quantize(xr, sc, zr, maxq)
                return ((qx - xr).abs().double() ** 2.4).sum(1)
            e_gpu, e_ref = _err(q.scale, q.zero), _err(s_ref, z_ref)
            differs = (q.scale.cpu().reshape(-1)[: xr.shape[0]] != s_ref.reshape(-1)[: xr.shape[0]])
            worst = float(((e_gpu - e_ref).abs() / e_ref.clamp_min(1e-30))[differs].max()) if differs.any() else 0.0
            _log(test="quantizer_mse", case=c["name"], rows_differing=int(differs.sum()), worst_rel_err_gap=worst)
            assert worst <= 1e-5, (c["name"], worst)
            continue
        assert torch.equal(q.scale.cpu(), s_ref), c["name"]
        assert torch.equal(q.zero.cpu(), z_ref), c["name"]
        xq = q.quantize(x)
        assert torch.equal(xq.cpu(), torch.from_numpy(z[c["name"] + "/xq"])), c["name"]


@pytest.mark.parametrize("gs", [-1, 32, 64, 128, 100])
@pytest.mark.parametrize("sym", [False, True])
def test_find_params_groups_vs_oracle(gs, sym):
    g = torch.Generator().manual_seed(3)
    W = torch.randn(70, 328, generator=g) * 0.03
    W[5, 64:128] = 0
    qp = O.QParams(bits=4, sym=sym)
    s, zr = Q.find_params_groups(W.to(DEV), qp.maxq, sym, groupsize=gs)
    step = gs if gs > 0 else 328
    for gi, c0 in enumerate(range(0, 328, step)):
        so, zo = O.find_params(W[:, c0:c0 + step], qp)
        assert torch.equal(s[:, gi].cpu(), so.flatten()), (gs, gi)
        assert torch.equal(zr[:, gi].cpu(), zo.flatten()), (gs, gi)


# ------------------------------------------------------------------------------------------ sgemm
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,Kd", [(128, 128, 128), (300, 200, 77), (513, 129, 260), (64, 1000, 128)])
def test_sgemm(ta, tb, M, N, Kd):
    g = torch.Generator().manual_seed(M + N + Kd)
    A = torch.randn((Kd, M) if ta else (M, Kd), generator=g)
    B = torch.randn((N, Kd) if tb else (Kd, N), generator=g)
    Cm = torch.randn(M, N, generator=g)
    want = -0.5 * ((A.t() if ta else A).double() @ (B.t() if tb else B).double()) + 2.0 * Cm.double()
    Ad, Bd, Cd = A.to(DEV), B.to(DEV), Cm.to(DEV).clone()
    _lib.check(_lib.load().gptq_b200_sgemm(ta, tb, M, N, Kd, -0.5, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0),
                                           2.0, Cd.data_ptr(), N, _lib.stream()), "sgemm")
    err = O.rel_fro(Cd.cpu(), want)
    _log(test="sgemm", ta=ta, tb=tb, M=M, N=N, K=Kd, rel=err)
    assert err < 2e-6


# ------------------------------------------------------------------------------------------ Cholesky chain
def _spd(K, seed, cond_boost=1.0):
    g = torch.Generator().manual_seed(seed)
    chan = torch.exp(0.5 * torch.randn(K, generator=g)) * cond_boost
    X = torch.randn(4 * K, K, generator=g) * chan
    H = (X.t().double() @ X.double()) * (2.0 / 4)
    return H.float()


def _hinv_upper_gpu(H, percdamp=0.01):
    lib = _lib.load()
    K = H.shape[0]
    Kp = (K + 127) // 128 * 128
    Hd = H.to(DEV).contiguous()
    diag = torch.empty(K, device=DEV)
    damp = torch.empty(1, device=DEV)
    _lib.check(lib.gptq_b200_hessian_prepare(Hd.data_ptr(), K, K, percdamp, diag.data_ptr(), damp.data_ptr(), 0, _lib.stream()), "prep")
    A = torch.empty((Kp, Kp), device=DEV)
    _lib.check(lib.gptq_b200_hessian_gather(Hd.data_ptr(), K, K, 0, damp.data_ptr(), A.data_ptr(), Kp, Kp, _lib.stream()), "gather")
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    n = lib.gptq_b200_hinv_upper_workspace_bytes(Kp)
    ws = torch.empty(n, dtype=torch.uint8, device=DEV)
    U = torch.empty((Kp, Kp), device=DEV)
    _lib.check(lib.gptq_b200_hinv_upper(A.data_ptr(), Kp, Kp, U.data_ptr(), Kp, info.data_ptr(), ws.data_ptr(), n, _lib.stream()), "hinv")
    torch.cuda.synchronize()
    return U[:K, :K].cpu(), int(info.item()), float(damp.item())


@pytest.mark.parametrize("K", [128, 200, 256, 640, 1024, 1408])
def test_hinv_upper_vs_fp64(K):
    H = _spd(K, K)
    U, info, damp = _hinv_upper_gpu(H)
    assert info == 0
    Hd = H.double().clone()
    Hd += torch.eye(K, dtype=torch.float64) * (0.01 * torch.mean(torch.diag(H)).double())
    L = torch.linalg.cholesky(Hd)
    U64 = torch.linalg.cholesky(torch.cholesky_inverse(L), upper=True)
    U32 = O.hinv_upper(H.clone(), 0.01)                       # the reference's own fp32 3-step route (CPU)
    e_gpu = O.rel_fro(U, U64)
    e_ref = O.rel_fro(U32, U64)
    _log(test="hinv_upper", K=K, rel_gpu_vs_fp64=e_gpu, rel_ref_vs_fp64=e_ref)
    assert torch.all(torch.tril(U, -1) == 0)
    assert e_gpu < max(5 * e_ref, 1e-5)


@pytest.mark.parametrize("K,panel", [(1024, 2), (1408, 0), (1408, 1), (2304, 4), (2176, 3), (4096, 0), (5120, 0), (7296, 0)])
def test_hinv_upper_lookahead_schedule(K, panel):
    """Right-looking panels with look-ahead on three internal streams (csrc/chol.cu hinv_lookahead; CPU restatement and
    race check of the schedule: oracle/chain_algorithms.py LookaheadSchedule) against the plain recursion and fp64:
    as close to fp64 as the recursion, and bitwise reproducible from run to run (every update is applied once, in an
    order fixed by the events, whatever the streams' interleaving)."""
    lib = _lib.load()
    H = _spd(K, K + panel)
    Hd = H.double().clone()
    Hd += torch.eye(K, dtype=torch.float64) * (0.01 * torch.mean(torch.diag(H)).double())
    U64 = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
    try:
        _lib.check(lib.gptq_b200_set_chain_schedule(0, -1), "set")
        U_rec, info_rec, _ = _hinv_upper_gpu(H)
        _lib.check(lib.gptq_b200_set_chain_schedule(1, panel), "set")
        runs = [_hinv_upper_gpu(H) for _ in range(3)]
    finally:
        _lib.check(lib.gptq_b200_set_chain_schedule(1, 0), "set")
    assert info_rec == 0 and all(r[1] == 0 for r in runs)
    U_la = runs[0][0]
    for r in runs[1:]:
        assert torch.equal(r[0], U_la)
    assert torch.all(torch.tril(U_la, -1) == 0)
    e_rec, e_la = O.rel_fro(U_rec, U64), O.rel_fro(U_la, U64)
    _log(test="hinv_upper_lookahead", K=K, panel=panel, rel_rec_vs_fp64=e_rec, rel_lookahead_vs_fp64=e_la,
         rel_lookahead_vs_rec=O.rel_fro(U_la, U_rec))
    assert e_la < max(3 * e_rec, 1e-6)


def test_hinv_upper_lookahead_not_pd_and_other_stream():
    """info is reported from the look-ahead schedule as well, and the call is ordered on a non-default caller stream."""
    lib = _lib.load()
    H = _spd(1280, 9)
    Hbad = H.clone()
    Hbad[700, 700] = -50.0
    _, info, _ = _hinv_upper_gpu(Hbad)
    assert info != 0
    U0, info0, _ = _hinv_upper_gpu(H)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        U1, info1, _ = _hinv_upper_gpu(H)
    assert info0 == 0 and info1 == 0 and torch.equal(U0, U1)


def test_hinv_upper_not_pd():
    H = _spd(256, 5)
    H[100, 100] = -50.0
    _, info, _ = _hinv_upper_gpu(H)
    assert info != 0


# ------------------------------------------------------------------------------------------ Hessian syrk
def _syrk(X, K, kind, alpha=1.0, beta=0.0, H0=None, chunk=0):
    lib = _lib.load()
    T = X.shape[0]
    H = torch.zeros((K, K), device=DEV) if H0 is None else H0.clone()
    if kind in ("tc", "tc1", "tc2", "tc3"):
        variant = {"tc": _lib.SYRK_AUTO, "tc1": _lib.SYRK_CTA1, "tc2": _lib.SYRK_CTA2, "tc3": _lib.SYRK_CTA2_BK128}[kind]
        n = lib.gptq_b200_syrk_workspace_bytes(K, T, variant)
        ws = torch.empty(max(n, 16), dtype=torch.uint8, device=DEV)
        _lib.check(lib.gptq_b200_syrk_accum(X.data_ptr(), _lib.dt(X), T, K, X.stride(0), H.data_ptr(), K, alpha, beta, chunk,
                                            variant, ws.data_ptr(), n, _lib.stream()), "syrk_accum")
    else:
        _lib.check(lib.gptq_b200_syrk_accum_simt(X.data_ptr(), _lib.dt(X), T, K, X.stride(0), H.data_ptr(), K, alpha, beta,
                                                 _lib.stream()), "syrk_simt")
    _lib.check(lib.gptq_b200_symmetrize(H.data_ptr(), K, K, _lib.stream()), "symmetrize")
    torch.cuda.synchronize()
    return H


def _acts(T, K, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    chan = torch.exp(0.5 * torch.randn(K, generator=g))
    return (torch.randn(T, K, generator=g) * chan).to(dtype)


@pytest.mark.parametrize("kind", ["tc1", "tc2", "tc3"])       # cta_group::1 (M128 x N256 per SM) and cta_group::2 (M256 x N256 per pair)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("T,K", [(64, 64), (100, 200), (2048, 256), (2048, 768), (8192, 1024), (4096, 2048), (333, 136),
                                 (1000, 520), (640, 2312)])
def test_syrk_tc_vs_fp64(T, K, dtype, kind):
    X = _acts(T, K, T + K, dtype)
    want = X.double().t() @ X.double()
    got = _syrk(X.to(DEV), K, kind)
    err = O.rel_fro(got.cpu(), want)
    sim = O.rel_fro(_syrk(X.to(DEV), K, "simt").cpu(), want)
    _log(test="syrk_tc", kind=kind, T=T, K=K, dtype=str(dtype), rel_tc=err, rel_simt=sim)
    assert sim < 1e-5
    assert err < 1e-5
    assert torch.equal(got, got.t())


def test_syrk_variants_agree_llama_shape():
    """Both tensor-core kernels on one staging window of the headline shape (16 samples x 2048 tokens, K = 4096): same H up
    to the summation order inside the tensor core, and a checksum property that does not need an fp64 K x K product:
    1^T H 1 = sum_t (sum_c x_tc)^2."""
    T, K = 32768, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    X = (torch.randn(T, K, device=DEV, generator=g) * torch.exp(0.5 * torch.randn(K, device=DEV, generator=g))).half()
    H1 = _syrk(X, K, "tc1")
    H2 = _syrk(X, K, "tc2")
    rel = float(torch.linalg.norm(H1 - H2) / torch.linalg.norm(H1))
    want = float((X.double().sum(1) ** 2).sum())
    got = float(H2.double().sum())
    tr = float((X.double() ** 2).sum())
    _log(test="syrk_variants", rel_1_vs_2=rel, checksum_rel=abs(got - want) / abs(want), trace_rel=abs(float(H2.diagonal().double().sum()) - tr) / tr)
    assert rel < 2e-6
    assert abs(got - want) <= 1e-5 * abs(tr)
    assert abs(float(H2.diagonal().double().sum()) - tr) <= 1e-5 * tr


def test_syrk_tc_alpha_beta_and_chunks():
    T, K = 4096, 512
    X = _acts(T, K, 11, torch.float16)
    H0 = torch.randn(K, K)
    H0 = (H0 + H0.t()).to(DEV)
    want = 0.5 * H0.cpu().double() + 0.25 * (X.double().t() @ X.double())
    for chunk in (0, 64, 512, 4096):
        for kind in ("tc1", "tc2", "tc3"):
            got = _syrk(X.to(DEV), K, kind, alpha=0.25, beta=0.5, H0=H0, chunk=chunk)
            err = O.rel_fro(got.cpu(), want)
            _log(test="syrk_tc_alpha_beta", kind=kind, chunk=chunk, rel=err)
            assert err < 1e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("grouped", [False, True])
def test_dequant_codes_rebuilds_the_written_weights(grouped, dtype):
    """Multi-GPU exchange #2 ships codes only: scale * (code - zero), cast to the layer dtype, must be bit-identical to the
    weights the column loop writes (quant.py:10, gptq.py:305)."""
    lib = _lib.load()
    R, K, gs = 70, 328, 41
    g = torch.Generator(device="cuda").manual_seed(5)
    ng = (K + gs - 1) // gs if grouped else 1
    codes = torch.randint(0, 16, (R, K), device=DEV, dtype=torch.uint8, generator=g)
    scale = torch.rand((R, ng), device=DEV, generator=g) * 0.02 + 0.003
    zero = torch.randint(0, 16, (R, ng), device=DEV, generator=g).float()
    gidx = (torch.randperm(K, device=DEV, generator=g) // gs).to(torch.int32) if grouped else torch.zeros(K, device=DEV, dtype=torch.int32)
    out = torch.empty((R, K), device=DEV, dtype=dtype)
    _lib.check(lib.gptq_b200_dequant_codes(codes.data_ptr(), K, R, K, scale.data_ptr(), zero.data_ptr(), ng, gidx.data_ptr(),
                                           out.data_ptr(), _lib.dt(out), K, _lib.stream()), "dequant_codes")
    want = (scale[:, gidx.long()] * (codes.float() - zero[:, gidx.long()])).to(dtype)
    assert torch.equal(out, want)


@pytest.mark.parametrize("K,ld", [(1, 4), (7, 8), (130, 136), (1000, 1000), (4096, 4096)])
def test_tril_pack_unpack_roundtrip(K, ld):
    """Multi-GPU exchange #1 carries the packed lower triangle (K (K + 1) / 2 floats): pack / unpack are exact inverses on
    the lower triangle and never touch the strict upper triangle."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(K)
    S = torch.randn((K, ld), device=DEV, generator=g)
    n = K * (K + 1) // 2
    packed = torch.full((n + 3,), -7.0, device=DEV)
    _lib.check(lib.gptq_b200_tril_pack(S.data_ptr(), K, ld, packed.data_ptr(), _lib.stream()), "tril_pack")
    idx = torch.tril_indices(K, K, device=DEV)
    assert torch.equal(packed[:n], S[idx[0], idx[1]])              # row-major over the lower triangle
    assert torch.all(packed[n:] == -7.0)
    out = torch.full((K, ld), 3.0, device=DEV)
    _lib.check(lib.gptq_b200_tril_unpack(packed.data_ptr(), K, out.data_ptr(), ld, _lib.stream()), "tril_unpack")
    want = torch.full((K, ld), 3.0, device=DEV)
    want[idx[0], idx[1]] = S[idx[0], idx[1]]
    assert torch.equal(out, want)


def test_syrk_accumulation_bias_long():
    # long reduction (32768 tokens, all-positive diagonal): exposes tensor-core accumulator rounding
    T, K = 32768, 256
    X = _acts(T, K, 5, torch.float16)
    want = X.double().t() @ X.double()
    for chunk in (512, 2048, 4096, 32768):
        got = _syrk(X.to(DEV), K, "tc", chunk=chunk)
        d = (torch.diag(got.cpu()).double() - torch.diag(want)) / torch.diag(want)
        _log(test="syrk_bias", chunk=chunk, rel_fro=O.rel_fro(got.cpu(), want), diag_rel_mean=float(d.mean()),
             diag_rel_max=float(d.abs().max()))
    assert O.rel_fro(_syrk(X.to(DEV), K, "tc").cpu(), want) < 1e-5


@pytest.mark.parametrize("tag", ["A", "B"])
def test_add_batch_matches_reference(tag, golden_A, golden_B):
    z, meta = golden_A if tag == "A" else golden_B
    lin = nn.Linear(meta["K"], meta["R"], bias=False).half().to(DEV)
    g = GPTQ(lin)
    for x in torch.from_numpy(z["X"]):
        g.add_batch(x.to(DEV), None)
    assert g.nsamples == meta["nsamples"]
    H = g.H.cpu()
    err = O.rel_fro(H, torch.from_numpy(z["H"]))
    _log(test="add_batch", tag=tag, rel_fro=err)
    assert err < 1e-5
    assert torch.equal(H, H.t())
    im = g.input.cpu()
    assert torch.allclose(im, torch.from_numpy(z["input_mean"]), rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------ fasterquant
def _golden_cases(tag):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"gptq_{tag}.npz"))
    return [(tag, c["name"]) for c in json.loads(bytes(z["meta"]).decode())["cases"]]


def _run_gptq(W16, batches, c, H_override=None):
    R, K = W16.shape
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    lin.weight.data = W16.to(DEV).clone()
    g = GPTQ(lin)
    g.quantizer = Q.Quantizer()
    g.quantizer.configure(c["bits"], perchannel=True, sym=c.get("sym", False), mse=c.get("mse", False), trits=c.get("trits", False))
    for x in batches:
        g.add_batch(x.to(DEV), None)
    if H_override is not None:
        g.H = H_override.to(DEV).clone()
    g.fasterquant(blocksize=128, percdamp=c.get("percdamp", .01), groupsize=c.get("groupsize", -1),
                  actorder=c.get("actorder", False), static_groups=c.get("static_groups", False))
    return lin, g


@pytest.mark.parametrize("tag,name", _golden_cases("A") + _golden_cases("B"))
def test_fasterquant_matches_reference(tag, name, golden_A, golden_B):
    z, meta = golden_A if tag == "A" else golden_B
    c = next(c for c in meta["cases"] if c["name"] == name)
    W16 = torch.from_numpy(z["W"])
    lin, g = _run_gptq(W16, list(torch.from_numpy(z["X"])), c)
    Qref = torch.from_numpy(z[f"{name}/Q"])
    Qgpu = lin.weight.data.cpu()
    mism = float((Qgpu != Qref).double().mean())
    rel_err = abs(g.error - c["error"]) / max(abs(c["error"]), 1e-30)
    _log(test="fasterquant_golden", tag=tag, case=name, mismatch=mism, err_gpu=g.error, err_ref=c["error"])
    budget = 1e-3 if not c.get("mse", False) else 5e-2
    assert mism <= budget
    assert rel_err <= 1e-2
    # codes dequantise to the written weights
    if not c.get("trits", False):
        s = g.scales[:, g.g_idx.long()]
        zr = g.zeros[:, g.g_idx.long()]
        deq = (s * (g.codes.float() - zr)).half()
        assert torch.equal(deq.cpu(), Qgpu)


@pytest.mark.parametrize("cfg", [
    dict(name="opt125m_b4", R=768, K=768, bits=4),
    dict(name="opt1.3b_b3_g128", R=512, K=2048, bits=3, groupsize=128),
    dict(name="llama_b4_g128_act_static", R=512, K=1024, bits=4, groupsize=128, actorder=True, static_groups=True),
    dict(name="bloom_b2_g64", R=384, K=1024, bits=2, groupsize=64),
    dict(name="opt66b_b3_act", R=256, K=1152, bits=3, actorder=True),
])
def test_fasterquant_vs_oracle_medium(cfg):
    """The five BASELINE.json configurations at reduced width; the oracle runs live on the host CPU.

    Solver parity is asserted on IDENTICAL H (the oracle's fp32 Hessian handed to the GPU solver);
    the end-to-end run (H from the tcgen05 syrk) is asserted against the same budget and logged."""
    R, K = cfg["R"], cfg["K"]
    g0 = torch.Generator().manual_seed(42)
    W16 = (torch.randn(R, K, generator=g0) * 0.02).half()
    # tokens >= 2K so that H is well conditioned without leaning on the damping term (T == K makes the codes
    # hypersensitive to 1e-6 perturbations of H, which is a property of the test matrix, not of the kernels)
    batches = [_acts(max(512, K // 2), K, 100 + i, torch.float16).unsqueeze(0) for i in range(4)]
    st = O.HessianState(K)
    for b in batches:
        O.add_batch(st, b)
    qp = O.QParams(bits=cfg["bits"])
    ref = O.fasterquant(W16, st.H, qp, groupsize=cfg.get("groupsize", -1), actorder=cfg.get("actorder", False),
                        static_groups=cfg.get("static_groups", False), keep_margin=True, stable_perm=True)
    for mode in ("same_H", "gpu_H"):
        lin, g = _run_gptq(W16, batches, cfg, H_override=st.H if mode == "same_H" else None)
        # rounding-tie classifier (oracle/gptq_oracle.py): a row may differ only from a column on where the oracle's
        # w/scale sat within TIE_TOL of a rounding boundary; everything outside such rows must be identical
        cls = O.classify_mismatches(g.codes, ref, TIE_TOL[mode], perm=getattr(g, "perm", None), diag=torch.diag(st.H))
        mism_q = float((lin.weight.data.cpu() != ref.Q.half()).double().mean())
        rel_err = abs(g.error - ref.error) / abs(ref.error)
        rec_gpu = O.recon_error(W16.float(), lin.weight.data.cpu().float(), st.H)
        rec_ref = O.recon_error(W16.float(), ref.Q, st.H)
        _log(test="fasterquant_medium", case=cfg["name"], H=mode, q_mismatch=mism_q, err_gpu=g.error, err_ref=ref.error,
             recon_gpu=rec_gpu, recon_ref=rec_ref, **cls)
        assert cls["unexplained_rows"] == 0 and cls["perm_unexplained_positions"] == 0, (mode, cls)
        # at this reduced width one tie row alone is > 0.1 % of the entries, so the 99.9 % gate is applied to the full
        # shapes in test_gpu_headline.py; here ties must stay rare
        assert cls["tie_rows"] <= max(2, R // 100), (mode, cls)
        assert rel_err <= 1e-2, mode
        assert abs(rec_gpu - rec_ref) <= 1e-2 * rec_ref, mode
        if cfg.get("actorder", False) and mode == "same_H":
            assert torch.equal(g.perm.cpu().long(), ref.perm)


def test_fasterquant_not_pd_raises():
    lin = nn.Linear(256, 16, bias=False).half().to(DEV)
    g = GPTQ(lin)
    g.quantizer = Q.Quantizer()
    g.quantizer.configure(4, perchannel=True, sym=False, mse=False)
    g.add_batch(_acts(64, 256, 1, torch.float16).to(DEV), None)
    H = g.H
    H[10, 10] = -1e6
    g.H = H
    with pytest.raises(torch.linalg.LinAlgError):
        g.fasterquant()


def test_research_modes_rejected_and_cpu_rejected():
    lin = nn.Linear(64, 8, bias=False).half().to(DEV)
    g = GPTQ(lin)
    g.quantizer = Q.Quantizer()
    g.quantizer.configure(4, perchannel=True, sym=False)
    with pytest.raises(NotImplementedError):
        g.fasterquant(lut_quant=True)
    with pytest.raises(_lib.GptqB200Error):
        GPTQ(nn.Linear(64, 8, bias=False))


# ------------------------------------------------------------------------------------------ tensor-core fp32 GEMM
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,Kd", [(128, 256, 32), (128, 128, 128), (300, 200, 96), (512, 640, 1024), (64, 1000, 132)])
def test_gemm3x_vs_fp64(ta, tb, M, N, Kd):
    """tcgen05 kind::tf32 3-pass split GEMM must be as accurate as the fp32 CUDA-core GEMM it replaces."""
    g = torch.Generator().manual_seed(7 * M + N + Kd)
    lda = ((M if ta else Kd) + 3) // 4 * 4
    ldb = ((Kd if tb else N) + 3) // 4 * 4
    A = torch.randn((Kd, lda) if ta else (M, lda), generator=g)
    B = torch.randn((N, ldb) if tb else (Kd, ldb), generator=g)
    Cm = torch.randn(M, N, generator=g)
    Aop = (A[:, :M].t() if ta else A[:, :Kd]).double()
    Bop = (B[:, :Kd].t() if tb else B[:, :N]).double()
    want = -0.5 * (Aop @ Bop) + 2.0 * Cm.double()
    Ad, Bd = A.to(DEV), B.to(DEV)
    lib = _lib.load()
    C1, C2 = Cm.to(DEV).clone(), Cm.to(DEV).clone()
    _lib.check(lib.gptq_b200_gemm3x(ta, tb, M, N, Kd, -0.5, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, 2.0, C1.data_ptr(), N,
                                    _lib.stream()), "gemm3x")
    _lib.check(lib.gptq_b200_sgemm(ta, tb, M, N, Kd, -0.5, Ad.data_ptr(), lda, Bd.data_ptr(), ldb, 2.0, C2.data_ptr(), N,
                                   _lib.stream()), "sgemm")
    torch.cuda.synchronize()
    e_tc, e_simt = O.rel_fro(C1.cpu(), want), O.rel_fro(C2.cpu(), want)
    _log(test="gemm3x", ta=ta, tb=tb, M=M, N=N, K=Kd, rel_tc=e_tc, rel_simt=e_simt)
    assert e_tc < 2e-6


def test_gemm3x_positive_sums_no_truncation_bias():
    """All-positive operands (the worst case for the truncating tensor-core accumulator), long K."""
    M, N, Kd = 128, 256, 4096
    g = torch.Generator().manual_seed(1)
    A = torch.rand(M, Kd, generator=g) + 0.5
    B = torch.rand(Kd, N, generator=g) + 0.5
    want = A.double() @ B.double()
    C = torch.zeros(M, N, device=DEV)
    Ad, Bd = A.to(DEV), B.to(DEV)
    _lib.check(_lib.load().gptq_b200_gemm3x(0, 0, M, N, Kd, 1.0, Ad.data_ptr(), Kd, Bd.data_ptr(), N, 0.0, C.data_ptr(), N,
                                            _lib.stream()), "gemm3x")
    torch.cuda.synchronize()
    rel = ((C.cpu().double() - want) / want)
    _log(test="gemm3x_bias", mean_rel=float(rel.mean()), max_rel=float(rel.abs().max()))
    assert float(rel.abs().max()) < 1.5e-6


@pytest.mark.parametrize("K", [256, 1408])
def test_hinv_upper_backends_agree(K):
    lib = _lib.load()
    H = _spd(K, K + 1)
    try:
        lib.gptq_b200_set_gemm_backend(0)
        U0, i0, _ = _hinv_upper_gpu(H)
        lib.gptq_b200_set_gemm_backend(1)
        U1, i1, _ = _hinv_upper_gpu(H)
    finally:
        lib.gptq_b200_set_gemm_backend(1)
    assert i0 == 0 and i1 == 0
    _log(test="hinv_backends", K=K, rel=O.rel_fro(U1, U0))
    assert O.rel_fro(U1, U0) < 2e-6


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_hoisted_reciprocal_division_is_exact(mode):
    """The column loop's a/b = f(a, b, RN(1/b)) must equal div.rn.f32 bit for bit (2^31 pairs per mode)."""
    out = torch.zeros(1, dtype=torch.int64, device=DEV)
    _lib.check(_lib.load().gptq_b200_selftest_div(1 << 31, 12345 + mode, mode, out.data_ptr(), _lib.stream()), "selftest_div")
    torch.cuda.synchronize()
    assert int(out.item()) == 0


# ------------------------------------------------------------------------------------------ packed matvec
@pytest.mark.parametrize("bits", [2, 3, 4, 8])
@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("xdtype", [torch.float32, torch.float16])
def test_quantlinear_forward_vs_oracle(bits, grouped, xdtype):
    R, K, gs = 200, 512, 64
    rng = np.random.default_rng(bits + 10 * grouped)
    ng = K // gs if grouped else 1
    codes = rng.integers(0, 2 ** bits, size=(R, K), dtype=np.uint8)
    scales = (rng.random((R, ng)) * 0.02 + 0.005).astype(np.float32)
    zeros = rng.integers(0, 2 ** bits, size=(R, ng)).astype(np.float32)
    g_idx = (rng.permutation(K) // gs).astype(np.int32) if grouped else None      # act-order style arbitrary groups
    bias = rng.standard_normal(R).astype(np.float32)
    x = rng.standard_normal(K).astype(np.float32)
    ql = Q.QuantLinear(bits, K, R)
    ql.qweight = Q.pack_codes(torch.from_numpy(codes).to(DEV), bits)
    ql.scales = torch.from_numpy(scales).to(DEV)
    ql.zeros = torch.from_numpy(zeros * scales).to(DEV)
    ql.bias = torch.from_numpy(bias).to(DEV)
    if grouped:
        ql.g_idx = torch.from_numpy(g_idx).to(DEV)
    xt = torch.from_numpy(x).to(DEV).to(xdtype)
    y = ql(xt.reshape(1, 1, K))
    assert y.shape == (1, 1, R) and y.dtype == xdtype
    want = O.vecquant_matvec(xt.float().cpu().numpy(), ql.qweight.cpu().numpy(), scales, zeros * scales, bits, bias, g_idx)
    tol = 2e-3 if xdtype == torch.float16 else 1e-4
    err = np.abs(y.float().cpu().numpy().reshape(-1) - want).max() / np.abs(want).max()
    _log(test="vecquant", bits=bits, grouped=grouped, dtype=str(xdtype), rel_max=float(err))
    assert err < tol
    with pytest.raises(ValueError):
        ql(torch.zeros(2, K, device=DEV))


def test_pack_then_forward_matches_dense_layer(golden_A):
    """Reference flow opt.py:362-373 + :440-507: fasterquant -> pack -> packed forward == dense forward of the quantized layer."""
    z, meta = golden_A
    c = next(c for c in meta["cases"] if c["name"] == "b3_perchannel")
    lin, g = _run_gptq(torch.from_numpy(z["W"]), list(torch.from_numpy(z["X"])), c)
    ql = Q.Quant3Linear(meta["K"], meta["R"])
    ql.pack(lin, g.quantizer.scale, g.quantizer.zero)
    x = torch.randn(1, meta["K"], device=DEV)
    y = ql(x)
    y_ref = x.double() @ lin.weight.data.double().t()
    assert torch.allclose(y.double(), y_ref, rtol=1e-3, atol=1e-4)
