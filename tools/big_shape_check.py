"""Sanity + timing at the wide shapes of the BASELINE configs that do not appear in the bench workload
(BLOOM-7b1 4h->h K=16384 2-bit g64; OPT-66B attention K=9216 3-bit act-order)."""
import contextlib, io, json, sys, time
import torch, torch.nn as nn
sys.path.insert(0, ".")
from gptq_b200.gptq import GPTQ
from gptq_b200.quant import Quantizer, pack_codes, unpack_codes
from oracle import gptq_oracle as O

DEV = torch.device("cuda:0")
for name, R, K, bits, kw in (("bloom7b1_4h_to_h", 2048, 16384, 2, dict(groupsize=64)),
                             ("opt66b_attn", 2304, 9216, 3, dict(actorder=True)),
                             ("opt66b_fc2_slice", 512, 36864, 3, dict(actorder=True))):
    g0 = torch.Generator(device="cuda").manual_seed(0)
    NS = 16 if K < 30000 else 20
    lin = nn.Linear(K, R, bias=False, device=DEV, dtype=torch.float16)
    lin.weight.data = (torch.randn(R, K, device=DEV, generator=g0) * 0.02).half()
    W0 = lin.weight.data.clone()
    chan = torch.exp(0.5 * torch.randn(K, device=DEV, generator=g0))
    g = GPTQ(lin)
    g.quantizer = Quantizer()
    g.quantizer.configure(bits, perchannel=True, sym=False, mse=False)
    torch.cuda.synchronize(); t0 = time.time()
    for j in range(NS):
        x = (torch.randn(1, 2048, K, device=DEV, generator=g0) * chan).half()
        g.add_batch(x, None)
    H = g.H
    torch.cuda.synchronize(); t_h = time.time() - t0
    # H check on a 512x512 corner against fp64
    g1 = torch.Generator(device="cuda").manual_seed(0)
    _ = torch.randn(R, K, device=DEV, generator=g1); _ = torch.randn(K, device=DEV, generator=g1)
    S = torch.zeros(512, 512, dtype=torch.float64, device=DEV)
    for j in range(NS):
        x = (torch.randn(1, 2048, K, device=DEV, generator=g1) * chan).half()[0, :, K - 512:].double()
        S += x.t() @ x
    S *= 2.0 / NS
    h_rel = float(torch.linalg.norm(H[K - 512:, K - 512:].double() - S) / torch.linalg.norm(S))
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        g.fasterquant(**kw)
    t_fq = time.time() - t0
    maxq = 2 ** bits - 1
    ok_codes = int(g.codes.max()) <= maxq
    s = g.scales[:, g.g_idx.long()]; z = g.zeros[:, g.g_idx.long()]
    deq_ok = torch.equal((s * (g.codes.float() - z)).half(), lin.weight.data)
    qw = pack_codes(g.codes, bits)
    rt_ok = torch.equal(unpack_codes(qw, bits), g.codes)
    rel_change = float((lin.weight.data.float() - W0.float()).norm() / W0.float().norm())
    print(json.dumps(dict(case=name, R=R, K=K, bits=bits, samples=NS, hessian_s=t_h, fasterquant_s=t_fq, H_corner_rel_fro=h_rel,
                          error=g.error, finite=bool(torch.isfinite(lin.weight.data).all()), codes_in_range=ok_codes,
                          codes_dequantise_to_weights=deq_ok, pack_roundtrip=rt_ok, rel_weight_change=rel_change,
                          mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
    g.free(); del g, lin, H
    torch.cuda.empty_cache()
