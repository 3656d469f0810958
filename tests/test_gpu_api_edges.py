"""Edge cases of the reference-facing API: layer kinds, input ranks, ragged / empty batches, odd sizes,
non-default block sizes -- every result checked against the CPU oracle."""
import contextlib
import copy
import io

import pytest
import torch
import torch.nn as nn

from oracle import gptq_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gptq_b200.gptq import GPTQ
    from gptq_b200 import quant as Q
    DEV = torch.device("cuda:0")


def _acts(T, K, seed, dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(T, K, generator=g) * torch.exp(0.5 * torch.randn(K, generator=g))).to(dtype)


def _gptq(layer, bits=4, **cfg):
    g = GPTQ(layer)
    g.quantizer = Q.Quantizer()
    g.quantizer.configure(bits, perchannel=True, sym=cfg.get("sym", False), mse=False)
    return g


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("R,K", [(1, 64), (7, 100), (33, 72), (130, 1000), (64, 8)])
def test_odd_shapes_match_oracle(R, K):
    W = (torch.randn(R, K, generator=torch.Generator().manual_seed(R + K)) * 0.02).half()
    X = [_acts(max(96, K) + 8 * i, K, i) for i in range(3)]              # ragged token counts, 2-D inputs, tokens >= 3K
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    lin.weight.data = W.to(DEV)
    g = _gptq(lin)
    st = O.HessianState(K)
    for x in X:
        g.add_batch(x.to(DEV), None)
        O.add_batch(st, x)
    assert g.nsamples == st.nsamples == 3
    assert O.rel_fro(g.H.cpu(), st.H) < 1e-5
    g.H = st.H.to(DEV).clone()                                            # identical H -> exact comparison of the solver
    _quiet(g.fasterquant, groupsize=32 if K >= 64 else -1)
    ref = O.fasterquant(W, st.H, O.QParams(bits=4), groupsize=32 if K >= 64 else -1, keep_margin=True)
    cls = O.classify_mismatches(g.codes, ref, 2.0 ** -18)                # identical H: rows may differ only from a rounding tie on
    bad_rows = cls["rows_bad"]
    assert cls["unexplained_rows"] == 0 and bad_rows <= 1, cls
    if bad_rows == 0:
        # identical codes; with dynamic groups the group scale is fitted to the lazily UPDATED weights (gptq.py:255), whose
        # last fp32 bits depend on the summation order of the update GEMM, so the dequantised fp16 value may differ by an ulp
        got, want = lin.weight.data.cpu().float(), ref.Q.half().float()
        assert torch.allclose(got, want, rtol=2 ** -9, atol=1e-7)
        assert (got == want).float().mean() >= 0.98


def test_batched_3d_input_and_empty_batch():
    K, R = 128, 16
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    g = _gptq(lin)
    st = O.HessianState(K)
    x = torch.stack([_acts(64, K, i) for i in range(4)])                  # [4, 64, K]: tmp = 4 sequences at once (gptq.py:44)
    g.add_batch(x.to(DEV), None)
    O.add_batch(st, x)
    g.add_batch(torch.empty(1, 0, K, dtype=torch.float16, device=DEV), None)   # empty batch still counts one sample
    O.add_batch(st, torch.empty(1, 0, K, dtype=torch.float16))
    assert g.nsamples == st.nsamples == 5
    assert O.rel_fro(g.H.cpu(), st.H) < 1e-5
    # reading H in the middle of calibration must not disturb the accumulation
    y = _acts(64, K, 9)
    g.add_batch(y.to(DEV), None)
    O.add_batch(st, y)
    assert O.rel_fro(g.H.cpu(), st.H) < 1e-5
    assert torch.allclose(g.input.cpu(), st.last_input_mean, rtol=1e-4, atol=1e-6)


def test_bf16_model_and_activations():
    K, R = 256, 48
    W = (torch.randn(R, K, generator=torch.Generator().manual_seed(0)) * 0.02).bfloat16()
    lin = nn.Linear(K, R, bias=False).bfloat16().to(DEV)
    lin.weight.data = W.to(DEV)
    g = _gptq(lin)
    st = O.HessianState(K)
    for i in range(3):
        x = _acts(128, K, i, torch.bfloat16)
        g.add_batch(x.to(DEV).unsqueeze(0), None)
        O.add_batch(st, x.unsqueeze(0))
    assert O.rel_fro(g.H.cpu(), st.H) < 1e-5
    g.H = st.H.to(DEV).clone()
    _quiet(g.fasterquant, groupsize=128, actorder=True, static_groups=True)
    ref = O.fasterquant(W, st.H, O.QParams(bits=4), groupsize=128, actorder=True, static_groups=True)
    assert lin.weight.dtype == torch.bfloat16
    assert torch.equal(g.codes.cpu(), ref.codes)
    assert torch.equal(lin.weight.data.cpu(), ref.Q.bfloat16())


def test_fp32_activations_use_exact_fp32_path():
    K, R = 96, 8
    lin = nn.Linear(K, R, bias=False).to(DEV)
    g = _gptq(lin)
    st = O.HessianState(K)
    for i in range(2):
        x = _acts(80, K, i, torch.float32)
        g.add_batch(x.to(DEV), None)
        O.add_batch(st, x)
    assert O.rel_fro(g.H.cpu(), st.H) < 2e-6


def test_conv1d_layer_is_transposed_like_the_reference():
    import transformers
    K, R = 128, 40
    conv = transformers.Conv1D(R, K)                                      # weight stored [K, R] (gptq.py:31-32)
    conv.weight.data = (torch.randn(K, R, generator=torch.Generator().manual_seed(1)) * 0.02)
    conv = conv.half().to(DEV)
    W = conv.weight.data.t().cpu()
    g = _gptq(conv)
    assert (g.rows, g.columns) == (R, K)
    st = O.HessianState(K)
    x = _acts(200, K, 3)
    g.add_batch(x.to(DEV).unsqueeze(0), None)
    O.add_batch(st, x.unsqueeze(0))
    g.H = st.H.to(DEV).clone()
    _quiet(g.fasterquant)
    ref = O.fasterquant(W, st.H, O.QParams(bits=4))
    assert conv.weight.shape == (K, R)
    assert torch.equal(conv.weight.data.cpu(), ref.Q.half().t())


def test_conv2d_layer_unfolds_patches():
    conv = nn.Conv2d(4, 6, kernel_size=3, padding=1, bias=False).half().to(DEV)
    g = _gptq(conv)
    assert (g.rows, g.columns) == (6, 36)
    x = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0)).half()
    g.add_batch(x.to(DEV), None)
    unfold = nn.Unfold(3, padding=1)
    cols = unfold(x.float()).permute([1, 0, 2]).flatten(1)                # gptq.py:50-58: [K, tokens]
    want = (2.0 / 2) * (cols.double() @ cols.double().t())
    assert O.rel_fro(g.H.cpu(), want) < 1e-5
    _quiet(g.fasterquant)
    assert conv.weight.shape == (6, 4, 3, 3) and conv.weight.dtype == torch.float16


@pytest.mark.parametrize("blocksize", [32, 64, 100])
def test_non_default_blocksize(blocksize):
    K, R = 256, 32
    W = (torch.randn(R, K, generator=torch.Generator().manual_seed(5)) * 0.02).half()
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    lin.weight.data = W.to(DEV)
    g = _gptq(lin, bits=3)
    st = O.HessianState(K)
    x = _acts(600, K, 1)
    g.add_batch(x.to(DEV), None)
    O.add_batch(st, x)
    g.H = st.H.to(DEV).clone()
    _quiet(g.fasterquant, blocksize=blocksize, groupsize=64)
    ref = O.fasterquant(W, st.H, O.QParams(bits=3), blocksize=blocksize, groupsize=64)
    agree = O.code_agreement(g.codes.cpu(), ref.codes)
    assert agree >= 0.999


def test_quantizer_survives_deepcopy_cpu_and_state_dict():
    q = Q.Quantizer()
    q.configure(4, perchannel=True, sym=False, mse=False)
    W = torch.randn(8, 64, device=DEV)
    q.find_params(W)
    q2 = copy.deepcopy(q).cpu()
    assert q2.scale.device.type == "cpu" and torch.equal(q2.scale, q.scale.cpu())
    assert set(q.state_dict()) == {"maxq", "scale", "zero"}
    assert bool(q.ready()) and bool(q.enabled())
    xq = q.quantize(W)
    assert torch.equal(xq.cpu(), O.quantize(W.cpu(), q2.scale, q2.zero, 15))


def test_rtn_baseline_matches_reference_semantics():
    """opt.py:292-300: find_params on the whole weight, then quantize()."""
    W = (torch.randn(40, 96, generator=torch.Generator().manual_seed(2)) * 0.05)
    q = Q.Quantizer()
    q.configure(3, perchannel=True, sym=True, mse=False)
    q.find_params(W.to(DEV), weight=True)
    got = Q.quantize(W.to(DEV), q.scale, q.zero, q.maxq)
    qp = O.QParams(bits=3, sym=True)
    s, z = O.find_params(W, qp)
    assert torch.equal(got.cpu(), O.quantize(W, s, z, qp.maxq))


def _three_linears(K, R):
    Ws = [(torch.randn(R, K, generator=torch.Generator().manual_seed(s)) * 0.02).half() for s in range(3)]
    lins = []
    for W in Ws:
        lin = nn.Linear(K, R, bias=False).half().to(DEV)
        lin.weight.data = W.to(DEV)
        lins.append(lin)
    return Ws, lins


def test_linears_seeing_the_same_tensor_share_one_hessian_and_one_chain():
    """q/k/v-style linears (their hooks receive the very same tensor) share the accumulator, the prologue and the
    inverse-Cholesky chain; results equal the unshared run bit for bit; different activations never share."""
    from gptq_b200 import _lib
    import gptq_b200.gptq as G
    K, R = 256, 24
    Ws, lins = _three_linears(K, R)
    X = [_acts(300, K, i).to(DEV) for i in range(3)]
    gs = [_gptq(l) for l in lins]                                   # all constructed first, like the reference drivers
    other = _acts(300, K, 7).to(DEV)
    for x in X:
        for g in gs[:2]:
            g.add_batch(x.unsqueeze(0), None)                       # same memory, like the hooks of q/k/v
    gs[2].add_batch(other, None)                                    # different activations
    assert gs[0]._acc is gs[1]._acc and gs[0]._acc.members == 2
    assert gs[2]._acc is not gs[0]._acc
    assert gs[0].nsamples == gs[1].nsamples == 3
    st = O.HessianState(K)
    for x in X:
        O.add_batch(st, x.cpu())
    assert O.rel_fro(gs[1].H.cpu(), st.H) < 1e-5                    # reading H does not disturb the sharing
    counts = []
    for g in gs:
        n0 = _lib.launch_count()
        _quiet(g.fasterquant, groupsize=128, actorder=True)
        counts.append(_lib.launch_count() - n0)
    assert counts[1] < counts[0] and counts[2] >= counts[0] - 2     # 2nd reuses the chain, 3rd recomputes
    for W, g in zip(Ws[:2], gs[:2]):
        ref = O.fasterquant(W, st.H, O.QParams(bits=4), groupsize=128, actorder=True)
        assert O.code_agreement(g.codes.cpu(), ref.codes) >= 0.999
    # with sharing off every linear accumulates and solves on its own -- and gives bit-identical results
    G.SHARE_H = False
    try:
        _, lins2 = _three_linears(K, R)
        gs2 = [_gptq(l) for l in lins2[:2]]
        for x in X:
            for g in gs2:
                g.add_batch(x.unsqueeze(0), None)
        assert gs2[0]._acc is not gs2[1]._acc
        for g in gs2:
            _quiet(g.fasterquant, groupsize=128, actorder=True)
        for a, b, la, lb in zip(gs[:2], gs2, lins[:2], lins2[:2]):
            assert torch.equal(a.codes, b.codes)
            assert torch.equal(la.weight.data, lb.weight.data)
            assert a.error == b.error
    finally:
        G.SHARE_H = True


def test_shared_accumulator_forks_when_the_batches_diverge():
    """Two linears share their first batches, then receive different tensors: each must end with its own exact H."""
    K, R = 128, 8
    _, lins = _three_linears(K, R)
    a, b = _gptq(lins[0]), _gptq(lins[1])
    X = [_acts(200, K, i).to(DEV) for i in range(4)]
    Ya = _acts(200, K, 50).to(DEV)
    Yb = _acts(200, K, 51).to(DEV)
    for x in X[:2]:
        a.add_batch(x, None)
        b.add_batch(x, None)
    assert a._acc is b._acc
    a.add_batch(Ya, None)
    b.add_batch(Yb, None)                                           # diverges one batch behind the leader
    assert a._acc is not b._acc
    a.add_batch(X[2], None)
    b.add_batch(X[3], None)
    for g, seq in ((a, X[:2] + [Ya, X[2]]), (b, X[:2] + [Yb, X[3]])):
        st = O.HessianState(K)
        for x in seq:
            O.add_batch(st, x.cpu())
        assert g.nsamples == 4
        assert O.rel_fro(g.H.cpu(), st.H) < 1e-5


def test_in_place_modification_between_sibling_hooks_is_detected():
    """Same memory, different content: the device-side compare flags it and fasterquant refuses to continue."""
    K, R = 128, 8
    _, lins = _three_linears(K, R)
    a, b = _gptq(lins[0]), _gptq(lins[1])
    x = _acts(256, K, 3).to(DEV)
    a.add_batch(x, None)
    x[5, 7] += 1.0                                                  # the caller mutates the tensor between the two hooks
    b.add_batch(x, None)
    assert a._acc is b._acc
    with pytest.raises(RuntimeError, match="GPTQ_B200_SHARE_H"):
        _quiet(a.fasterquant)


def test_sibling_can_read_H_and_free_in_any_order():
    K, R = 128, 8
    Ws, lins = _three_linears(K, R)
    a, b = _gptq(lins[0]), _gptq(lins[1])
    X = [_acts(256, K, i).to(DEV) for i in range(2)]
    for x in X:
        a.add_batch(x, None)
        b.add_batch(x, None)
    st = O.HessianState(K)
    for x in X:
        O.add_batch(st, x.cpu())
    _quiet(a.fasterquant)
    a.free()
    assert O.rel_fro(b.H.cpu(), st.H) < 1e-5                        # b still sees the (finished) shared Hessian
    _quiet(b.fasterquant)                                           # uses the materialised copy: private solve
    ref = O.fasterquant(Ws[1], st.H, O.QParams(bits=4))
    assert O.code_agreement(b.codes.cpu(), ref.codes) >= 0.999
    with pytest.raises(RuntimeError):
        b.add_batch(X[0], None)                                     # consumed, like `del self.H` in the reference
    b.free()


def test_non_finite_activations_raise_linalgerror_instead_of_faulting():
    """A NaN in the calibration data poisons diag(H): the act-order argsort must still emit a permutation and the
    chain must report non-PD (torch.linalg.cholesky raises LinAlgError for the reference too, gptq.py:177)."""
    K, R = 256, 16
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    x = _acts(300, K, 0)
    x[3, 5] = float("nan")
    g = _gptq(lin)
    g.add_batch(x.to(DEV), None)
    with pytest.raises(torch.linalg.LinAlgError):
        _quiet(g.fasterquant, actorder=True, groupsize=128)
    torch.cuda.synchronize()                                          # the context is still healthy
    g2 = _gptq(nn.Linear(K, R, bias=False).half().to(DEV))
    g2.add_batch(_acts(300, K, 1).to(DEV), None)
    _quiet(g2.fasterquant, actorder=True, groupsize=128)


def test_siblings_are_solved_in_one_column_loop_and_match_individual_solves():
    """q/k/v-style linears: the first fasterquant solves the siblings' rows in the same column loop; each sibling's own
    call then only publishes.  Results must equal the ungrouped run (same codes up to rounding ties, same error)."""
    from gptq_b200 import _lib
    import gptq_b200.gptq as G
    K, R = 256, 40
    Ws, lins = _three_linears(K, R)
    X = [_acts(300, K, i).to(DEV) for i in range(2)]

    def run(group):
        G.GROUP_SIBLINGS = group
        try:
            _, ls = _three_linears(K, R)
            gs = [_gptq(l) for l in ls]
            for x in X:
                for g in gs:
                    g.add_batch(x, None)
            counts = []
            for g in gs:
                n0 = _lib.launch_count()
                _quiet(g.fasterquant, groupsize=64, actorder=True, static_groups=True)
                counts.append(_lib.launch_count() - n0)
            return gs, ls, counts
        finally:
            G.GROUP_SIBLINGS = True
    ga, la, ca = run(True)
    gb, lb, cb = run(False)
    assert ca[1] < cb[1] // 4 and ca[2] < cb[2] // 4                # siblings only publish
    for a, b, x, y in zip(ga, gb, la, lb):
        assert O.code_agreement(a.codes.cpu(), b.codes.cpu()) >= 0.999
        assert abs(a.error - b.error) <= 1e-3 * abs(b.error)
        assert torch.equal(a.scales, b.scales) and torch.equal(a.zeros, b.zeros)
        assert torch.equal(a.quantizer.scale, b.quantizer.scale)
        assert (x.weight.data == y.weight.data).float().mean() >= 0.999
    # a sibling whose weights changed after it was solved ahead is solved again with the new weights
    _, ls = _three_linears(K, R)
    gs = [_gptq(l) for l in ls]
    for x in X:
        for g in gs:
            g.add_batch(x, None)
    _quiet(gs[0].fasterquant, groupsize=64)
    ls[1].weight.data = (ls[1].weight.data * 1.5)
    st = O.HessianState(K)
    for x in X:
        O.add_batch(st, x.cpu())
    _quiet(gs[1].fasterquant, groupsize=64)
    ref = O.fasterquant((Ws[1].float() * 1.5).half(), st.H, O.QParams(bits=4), groupsize=64)
    assert O.code_agreement(gs[1].codes.cpu(), ref.codes) >= 0.999
    # ... and one called with other arguments as well
    _quiet(gs[2].fasterquant, groupsize=-1)
    ref = O.fasterquant(Ws[2], st.H, O.QParams(bits=4), groupsize=-1)
    assert O.code_agreement(gs[2].codes.cpu(), ref.codes) >= 0.999


def test_overlapped_syrk_on_side_stream_gives_the_same_hessian():
    """GPTQ_B200_OVERLAP_SYRK: the accumulation runs on a side stream with two staging slabs; same H, bit for bit."""
    import gptq_b200.gptq as G
    K, R = 256, 16
    X = [_acts(300, K, i).to(DEV) for i in range(40)]               # > 2 slabs of 16 samples
    Hs = []
    for flag in (False, True):
        G.OVERLAP_SYRK = flag
        try:
            g = _gptq(nn.Linear(K, R, bias=False).half().to(DEV))
            for x in X:
                g.add_batch(x, None)
            Hs.append(g.H.clone())
            _quiet(g.fasterquant, groupsize=128)
        finally:
            G.OVERLAP_SYRK = False
    torch.cuda.synchronize()
    assert torch.equal(Hs[0], Hs[1])
