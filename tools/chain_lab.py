"""A/B timing of the two schedules of gptq_b200_hinv_upper (plain recursion vs right-looking panels with look-ahead).

    python tools/chain_lab.py [K ...]          # default 4096 11008

Only the hinv_upper call is inside the CUDA-event bracket; the matrix is restored before every call.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gptq_b200 import _lib

DEV = torch.device("cuda", 0)
lib = _lib.load()


def run(K, lookahead, panel, iters=5):
    _lib.check(lib.gptq_b200_set_chain_schedule(lookahead, panel), "set")
    Kp = (K + 127) // 128 * 128
    g = torch.Generator(device=DEV).manual_seed(K)
    X = torch.randn(2 * K, K, device=DEV, generator=g)
    H = (X.t() @ X) / K + torch.eye(K, device=DEV)
    del X
    n = lib.gptq_b200_hinv_upper_workspace_bytes(Kp)
    ws = torch.empty(n, dtype=torch.uint8, device=DEV)
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    A = torch.empty(Kp, Kp, device=DEV)
    U = torch.empty(Kp, Kp, device=DEV)
    times = []
    for it in range(iters + 2):
        A.zero_()
        A[:K, :K].copy_(H)
        if Kp > K:
            A[K:, K:].copy_(torch.eye(Kp - K, device=DEV))
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.gptq_b200_hinv_upper(A.data_ptr(), Kp, Kp, U.data_ptr(), Kp, info.data_ptr(), ws.data_ptr(), n, _lib.stream()), "hinv")
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(a.elapsed_time(b))
    # residual of the defining identity  H^-1 = U^T U  on a random vector (fp64 on the device)
    v = torch.randn(K, 1, device=DEV, dtype=torch.float64, generator=torch.Generator(device=DEV).manual_seed(1))
    Ud = U[:K, :K].double()
    r = (Ud.t() @ (Ud @ (H.double() @ v)) - v).norm() / v.norm()
    return dict(op="hinv_upper", K=K, lookahead=lookahead, panel=panel, skip=os.environ.get("GPTQ_B200_CHAIN_SKIP", "0"), ms_min=min(times), ms_avg=sum(times) / len(times),
                info=int(info.item()), identity_residual=float(r)), U


if __name__ == "__main__":
    Ks = [int(a) for a in sys.argv[1:]] or [4096, 11008]
    for K in Ks:
        base, U0 = run(K, 0, 2)
        print(json.dumps(base), flush=True)
        for panel in (1, 2, 4):
            r, U1 = run(K, 1, panel)
            r["rel_diff_vs_recursion"] = float((U1 - U0).norm() / U0.norm())
            print(json.dumps(r), flush=True)
            del U1
    lib.gptq_b200_set_chain_schedule(1, 0)
