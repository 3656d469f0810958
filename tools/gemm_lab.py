"""3xTF32 tensor-core GEMM (gptq_b200_gemm3x) at the shapes of the lazy block updates and the chain's trailing updates,
C <- C - A B  (beta = 1), rotating over several C buffers so that C comes from HBM like in the solver.

    python tools/gemm_lab.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gptq_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)
lib = _lib.load()
SHAPES = [(4096, 3584, 512), (4096, 10496, 512), (11008, 3584, 512), (4096, 384, 128), (9984, 9984, 1024), (4096, 4096, 4096)]
for M, N, Kd in SHAPES:
    A = torch.randn(M, Kd, device=DEV)
    B = torch.randn(Kd, N, device=DEV)
    nbuf = max(2, int(400e6 // (M * N * 4)) + 1)
    Cs = [torch.zeros(M, N, device=DEV) for _ in range(min(nbuf, 8))]
    def run(i):
        C = Cs[i % len(Cs)]
        _lib.check(lib.gptq_b200_gemm3x(0, 0, M, N, Kd, -1.0, A.data_ptr(), Kd, B.data_ptr(), N, 1.0, C.data_ptr(), N, _lib.stream()), "gemm3x")
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    iters = 20
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        run(i)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(json.dumps(dict(op="gemm3x_beta1", M=M, N=N, K=Kd, ms=round(ms, 4), tflops=round(2.0 * M * N * Kd / ms / 1e9, 1),
                          cprefetch=os.environ.get("GPTQ_B200_GEMM_CPREFETCH", "default"))), flush=True)
    del A, B, Cs
