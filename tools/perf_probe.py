"""Per-kernel timing probe (CUDA events on the launching stream).  Usage: python tools/perf_probe.py [what ...]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from gptq_b200 import _lib
from gptq_b200 import quant as Q

DEV = torch.device("cuda:0")
lib = _lib.load()


ONCE = False          # "once" on the command line: every probe launches its kernel twice and skips the timing loop (ncu runs)


def timeit(fn, warm=2, iters=5):
    if ONCE:
        fn(); fn()
        torch.cuda.synchronize()
        return 1.0, 1.0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def probe_syrk(K, T, chunk=0, dtype=torch.float16, variant=0, iters=5):
    X = (torch.randn(T, K, device=DEV) * 0.5).to(dtype)
    H = torch.zeros(K, K, device=DEV)
    n = lib.gptq_b200_syrk_workspace_bytes(K, T, variant)
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_syrk_accum(X.data_ptr(), _lib.dt(X), T, K, K, H.data_ptr(), K, 1.0, 1.0, chunk, variant,
                                            ws.data_ptr(), n, _lib.stream()), "syrk")
    best, avg = timeit(run, iters=iters)
    nbm, nbn = (K + 127) // 128, (K + 255) // 256
    tiles = sum(nbm - 2 * j for j in range(nbn))
    issued = 2.0 * tiles * 128 * 256 * T          # = 256x256 pair tiles with bi >= bj: the same count for both kernels
    dense = 2.0 * K * K * T
    print(json.dumps(dict(op="syrk", variant=variant, K=K, T=T, chunk=chunk, ms=best, ms_avg=avg, issued_tflops=issued / best / 1e9,
                          dense_equiv_tflops=dense / best / 1e9)))


def probe_hinv(K):
    Kp = (K + 127) // 128 * 128
    X = torch.randn(2 * K, K, device=DEV)
    H = (X.t() @ X) / K + torch.eye(K, device=DEV)
    n = lib.gptq_b200_hinv_upper_workspace_bytes(Kp)
    ws = torch.empty(n, dtype=torch.uint8, device=DEV)
    info = torch.zeros(1, dtype=torch.int32, device=DEV)
    A = torch.empty(Kp, Kp, device=DEV)
    Uo = torch.empty(Kp, Kp, device=DEV)

    def run():
        A.zero_()
        A[:K, :K].copy_(H)
        if Kp > K:
            A[K:, K:].copy_(torch.eye(Kp - K, device=DEV))
        _lib.check(lib.gptq_b200_hinv_upper(A.data_ptr(), Kp, Kp, Uo.data_ptr(), Kp, info.data_ptr(), ws.data_ptr(), n, _lib.stream()), "hinv")
    best, avg = timeit(run, warm=1, iters=3)
    print(json.dumps(dict(op="hinv_upper", K=K, ms=best, gflops=(2.0 / 3 * K ** 3) / best / 1e6, info=int(info.item()))))
    t0 = time.time()
    Hc = H.clone()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    L = torch.linalg.cholesky(Hc)
    Hi = torch.cholesky_inverse(L)
    U = torch.linalg.cholesky(Hi, upper=True)
    b.record(); torch.cuda.synchronize()
    print(json.dumps(dict(op="torch_3step_chain_cuda", K=K, ms=a.elapsed_time(b))))


def probe_colloop(R, K, gs=128):
    Kp = (K + 127) // 128 * 128
    X = torch.randn(2 * K, K, device=DEV)
    H = (X.t() @ X) / K + torch.eye(K, device=DEV)
    U = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H)), upper=True).contiguous()
    W0 = torch.randn(R, K, device=DEV) * 0.02
    ng = (K + gs - 1) // gs
    scale, zero = Q.find_params_groups(W0, 15, False, groupsize=gs)
    gidx = (torch.arange(K, device=DEV) // gs).to(torch.int32)
    Qo = torch.empty(R, K, device=DEV)
    codes = torch.empty(R, K, dtype=torch.uint8, device=DEV)
    loss = torch.zeros(R, dtype=torch.float32, device=DEV)
    n = lib.gptq_b200_colloop_workspace_bytes(R, K, 128)
    ws = torch.empty(n, dtype=torch.uint8, device=DEV)
    W = torch.empty_like(W0)

    def run():
        W.copy_(W0)
        _lib.check(lib.gptq_b200_colloop(W.data_ptr(), R, K, K, U.data_ptr(), K, 128, gs, 0, gidx.data_ptr(), scale.data_ptr(),
                                         zero.data_ptr(), ng, 15, 0, 0, 2.4, 100, 0.8, Qo.data_ptr(), K, codes.data_ptr(), K,
                                         loss.data_ptr(), ws.data_ptr(), n, _lib.stream()), "colloop")
    best, avg = timeit(run, warm=1, iters=3)
    print(json.dumps(dict(op="colloop", R=R, K=K, ms=best, us_per_column=best * 1e3 / K, lazy_gemm_gflop=R * K * K / 1e9)))


def probe_pack(R, K, bits):
    codes = torch.randint(0, 2 ** bits, (R, K), dtype=torch.uint8, device=DEV)
    qw = torch.empty((K * bits // 32, R), dtype=torch.int32, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_pack_codes(codes.data_ptr(), R, K, K, bits, qw.data_ptr(), _lib.stream()), "pack")
    best, avg = timeit(run, warm=3, iters=10)
    byts = R * K * (1 + bits / 8)
    print(json.dumps(dict(op="pack_codes", R=R, K=K, bits=bits, us=best * 1e3, gbs=byts / best / 1e6)))


def probe_pack_weights(R, K, bits, gs=128):
    """The reference's entry point (quant.py:152-187): fp16 dequantised weights + scales / zeros in, packed words out."""
    ng = K // gs
    scale = torch.rand(R, ng, device=DEV) * 0.01 + 0.005
    zero = torch.randint(0, 2 ** bits, (R, ng), device=DEV).float()
    gidx = (torch.arange(K, device=DEV) // gs).to(torch.int32)
    codes = torch.randint(0, 2 ** bits, (R, K), device=DEV).float()
    W = (scale[:, gidx.long()] * (codes - zero[:, gidx.long()])).half()
    qw = torch.empty((K * bits // 32, R), dtype=torch.int32, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_pack_weights(W.data_ptr(), _lib.dt(W), R, K, K, scale.data_ptr(), zero.data_ptr(), ng,
                                              gidx.data_ptr(), bits, qw.data_ptr(), _lib.stream()), "pack_weights")
    best, _ = timeit(run, warm=3, iters=10)
    byts = R * K * (2 + bits / 8) + 8 * R * ng
    print(json.dumps(dict(op="pack_weights", R=R, K=K, bits=bits, us=best * 1e3, gbs=byts / best / 1e6)))


def probe_find_params(R, K, gs):
    W = torch.randn(R, K, device=DEV) * 0.02
    ng = 1 if gs == -1 else K // gs
    scale = torch.empty(R, ng, device=DEV)
    zero = torch.empty(R, ng, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_find_params(W.data_ptr(), R, K, K, 0, K, gs if gs > 0 else K, 15, 0, 0, 2.4, 100, 0.8,
                                             scale.data_ptr(), zero.data_ptr(), ng, 0, _lib.stream()), "find_params")
    best, _ = timeit(run, warm=3, iters=10)
    byts = 4.0 * R * K + 8.0 * R * ng
    print(json.dumps(dict(op="find_params", R=R, K=K, groupsize=gs, us=best * 1e3, gbs=byts / best / 1e6)))


def probe_vecquant(R, K, bits, gs=-1):
    ng = 1 if gs == -1 else K // gs
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K * bits // 32, R), dtype=torch.int32, device=DEV)
    scale = torch.rand(R, ng, device=DEV) * 0.01 + 0.005
    zs = torch.rand(R, ng, device=DEV)
    gidx = (torch.arange(K, device=DEV) // max(gs, 1)).to(torch.int32) if gs > 0 else None
    x = torch.randn(K, device=DEV).half()
    y = torch.zeros(R, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_vecquant_matvec(x.data_ptr(), _lib.dt(x), qw.data_ptr(), R, K, bits, scale.data_ptr(), zs.data_ptr(),
                                                 ng, _lib.ptr(gidx), 0, y.data_ptr(), _lib.stream()), "vecquant")
    best, _ = timeit(run, warm=3, iters=10)
    byts = R * K * bits / 8 + 2.0 * K + 4.0 * R + 8.0 * R * ng
    print(json.dumps(dict(op="vecquant_matvec", R=R, K=K, bits=bits, groupsize=gs, us=best * 1e3, gbs=byts / best / 1e6)))


def probe_sgemm(M, N, Kd):
    A = torch.randn(M, Kd, device=DEV); B = torch.randn(Kd, N, device=DEV); C = torch.zeros(M, N, device=DEV)
    for name, fn in (("sgemm", lib.gptq_b200_sgemm), ("gemm3x", lib.gptq_b200_gemm3x)):
        def run():
            _lib.check(fn(0, 0, M, N, Kd, 1.0, A.data_ptr(), Kd, B.data_ptr(), N, 0.0, C.data_ptr(), N, _lib.stream()), name)
        best, _ = timeit(run)
        print(json.dumps(dict(op=name, M=M, N=N, K=Kd, ms=best, tflops=2.0 * M * N * Kd / best / 1e9)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["syrk", "hinv", "colloop", "pack", "sgemm"]
    if "once" in what:
        ONCE = True
        what.remove("once")
    if "syrk" in what:
        for variant in (1, 2):
            for K, T in ((4096, 32768), (4096, 2048), (11008, 32768), (768, 32768), (2048, 32768)):
                probe_syrk(K, T, variant=variant)
        for chunk in (64, 128, 256, 1024, 4096):
            probe_syrk(4096, 32768, chunk)
    if "syrk2" in what:          # the two headline launch shapes, both kernels, interleaved (same clocks / power state)
        for rep in range(2):
            for K in (4096, 11008):
                for variant in (1, 2):
                    probe_syrk(K, 32768, variant=variant, iters=10)
        for chunk in (256, 1024, 2048):
            probe_syrk(4096, 32768, chunk, variant=2, iters=10)
    if "sgemm" in what:
        probe_sgemm(4096, 4096, 4096)
        probe_sgemm(4096, 4096, 128)
        probe_sgemm(128, 128, 128)
        probe_sgemm(512, 512, 512)
    if "diag" in what:
        probe_hinv(128)
        probe_hinv(256)
    if "hinv" in what:
        for K in (1024, 4096):
            probe_hinv(K)
    if "hinvbig" in what:
        probe_hinv(11008)
    if "colloop" in what:
        probe_colloop(4096, 4096)
        probe_colloop(11008, 4096)
    if "pack" in what:
        for bits in (2, 3, 4):
            probe_pack(4096, 11008, bits)
    if "hbm" in what:          # the HBM-bound byte kernels at Llama-2-7B down_proj / gate_proj shapes (ncu: achieved GB/s)
        probe_pack(4096, 11008, 4)
        probe_pack(4096, 11008, 3)
        probe_pack_weights(4096, 11008, 4)
        probe_pack_weights(4096, 11008, 3, gs=11008)
        probe_find_params(4096, 11008, 128)
        probe_find_params(11008, 4096, -1)
        probe_vecquant(11008, 4096, 3)
        probe_vecquant(11008, 4096, 4, gs=128)
        probe_vecquant(4096, 11008, 4, gs=128)
