"""Which stage makes the CUDA solver diverge from the oracle at non-tie columns?  (headline shape, identical H)

Runs the oracle once (keep_U, keep_margin) and then the CUDA column loop with (a) its own chain, (b) the ORACLE's upper
factor, each with the tensor-core (3xTF32) and the exact-FMA CUDA-core GEMM backend, and classifies the differing rows.
"""
import contextlib
import io
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import gptq_oracle as O
from gptq_b200 import _lib
import gptq_b200.gptq as G
from gptq_b200.quant import Quantizer

DEV = torch.device("cuda:0")
lib = _lib.load()


def acts(T, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    chan = torch.exp(0.5 * torch.randn(K, device=DEV, generator=g))
    return (torch.randn(T, K, device=DEV, generator=g) * chan).half()


def main():
    R, K, ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 4096, 8
    gw = torch.Generator().manual_seed(7)
    W16 = (torch.randn(4096, K, generator=gw) * 0.02).half()[:R]
    batches = [acts(2048, K, 1000 + j) for j in range(ns)]
    st = O.HessianState(K, H=torch.zeros((K, K), device=DEV))
    for b in batches:
        O.add_batch(st, b)
    H = st.H.cpu()
    kw = dict(groupsize=128, actorder=True, static_groups=True)
    ref = O.fasterquant(W16, H, O.QParams(bits=4), keep_U=True, keep_margin=True, **kw)
    q = Quantizer()
    q.configure(4, perchannel=True, sym=False, mse=False)
    qcfg = G.GPTQ._qcfg(q)

    def solve(U_override, backend):
        lib.gptq_b200_set_gemm_backend(backend)
        try:
            Hd = H.to(DEV).clone()
            sol = G._solve_hessian(Hd, K, 0.01, True)
            pg = sol["perm"].cpu().long()
            bad = torch.nonzero(pg != ref.perm).flatten()
            if len(bad):
                d = torch.diag(H)
                print(json.dumps(dict(perm_positions_differing=len(bad), first=bad[:6].tolist(),
                                      diag_gpu_order=[float(d[pg[i]]) for i in bad[:6]], diag_ref_order=[float(d[ref.perm[i]]) for i in bad[:6]],
                                      diag_is_sorted_gpu=bool((d[pg][:-1] >= d[pg][1:]).all()), diag_is_sorted_ref=bool((d[ref.perm][:-1] >= d[ref.perm][1:]).all()))), flush=True)
            if U_override is not None:
                sol["perm"] = ref.perm.to(torch.int32).to(DEV)
                sol["invperm"] = torch.argsort(ref.perm).to(torch.int32).to(DEV)
                sol["U"] = U_override.to(DEV).contiguous()
            with contextlib.redirect_stdout(io.StringIO()):
                out = G._solve_rows([W16.to(DEV)], sol, qcfg, 128, 128, True, True, torch.float16)[0]
            torch.cuda.synchronize()
            return out["codes"], sol["U"]
        finally:
            lib.gptq_b200_set_gemm_backend(1)

    for name, Uo, be in (("gpu_chain+tc_gemm", None, 1), ("oracle_U+tc_gemm", ref.U, 1), ("oracle_U+simt_gemm", ref.U, 0),
                         ("gpu_chain+simt_gemm(all simt)", None, 0)):
        codes, U = solve(Uo, be)
        cls = O.classify_mismatches(codes, ref, 2.0 ** -7)
        urel = O.rel_fro(U[:K, :K].cpu(), ref.U)
        print(json.dumps(dict(experiment=name, U_rel_fro_vs_oracle=urel, **cls)), flush=True)


if __name__ == "__main__":
    main()
