"""Hessian-kernel laboratory: sustained throughput with clocks / power sampled during the run, for each kernel variant and
for a cuBLAS bf16 GEMM of comparable size (the roofline denominator's own workload) in the SAME process and power state.

    python tools/syrk_lab.py sustained [seconds]     # JSON lines
    python tools/syrk_lab.py once                    # one launch per variant and shape (for ncu)
"""
import json
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from gptq_b200 import _lib

DEV = torch.device("cuda:0")
lib = _lib.load()


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        sm, pw, cap = [], [], 0
        for t, line in self.rows:
            if t0 <= t <= t1:
                c = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(c[0])); pw.append(float(c[1])); cap += c[2].lower().startswith("active")
                except Exception:
                    pass
        sm.sort(); pw.sort()
        return dict(samples=len(sm), sm_mhz_median=sm[len(sm) // 2] if sm else None, sm_mhz_min=sm[0] if sm else None,
                    power_w_median=pw[len(pw) // 2] if pw else None, power_cap_samples=cap)

    def stop(self):
        self.p.terminate()


def issued_flops(K, T):
    nb = (K + 255) // 256
    return 2.0 * (nb * (nb + 1) // 2) * 256 * 256 * T


def make_syrk(K, T, variant, chunk=0):
    X = (torch.randn(T, K, device=DEV) * 0.5).half()
    H = torch.zeros(K, K, device=DEV)
    n = lib.gptq_b200_syrk_workspace_bytes(K, T, variant)
    ws = torch.empty(max(n, 16), dtype=torch.uint8, device=DEV)

    def run():
        _lib.check(lib.gptq_b200_syrk_accum(X.data_ptr(), _lib.dt(X), T, K, K, H.data_ptr(), K, 1.0, 1.0, chunk, variant,
                                            ws.data_ptr(), n, _lib.stream()), "syrk")
    run.keep = (X, H, ws)
    return run


def sustained(fn, flops, seconds, smi, tag):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    n = 0
    a.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = a.elapsed_time(b) / n
    out = dict(tag=tag, launches=n, ms_per_launch=ms, tflops=flops / ms / 1e9)
    out.update(smi.window(t0 + 0.3, t1))
    print(json.dumps(out), flush=True)
    time.sleep(1.0)            # let the power state relax between experiments


SWEEP = [(8, 9, 0), (8, 9, 1), (6, 12, 0), (12, 6, 0), (4, 18, 0), (18, 4, 0), (5, 15, 0), (3, 24, 0), (43, 2, 0), (2, 43, 0), (10, 7, 0)]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "sustained"
    if what == "once":          # every launch shape of the default bench, default kernel, 2 launches each (ncu: use the 2nd)
        order = []
        for K, T in ((4096, 4096), (4096, 8192), (4096, 16384), (4096, 32768), (11008, 32768)):
            fn = make_syrk(K, T, 0)
            for _ in range(2):
                fn()
                order.append([K, T])
            torch.cuda.synchronize()
        with open("gpurun_out/r2_syrk_once_order.json", "w") as f:
            json.dump(order, f)
        return
    if what == "sweep":          # tile-order / TMA-promotion sweep of the pair kernel at K = 11008 (run under ncu for DRAM bytes)
        for sw, sh, p128 in SWEEP:
            v = 3 | (sw << 8) | (sh << 16) | (p128 << 24)
            fn = make_syrk(11008, 32768, v)
            fn(); fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            print(json.dumps(dict(sw=sw, sh=sh, p128=p128, ms=a.elapsed_time(b), tflops=issued_flops(11008, 32768) / a.elapsed_time(b) / 1e9)), flush=True)
        return
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
    variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2]
    smi = Smi()
    time.sleep(0.5)
    A = torch.randn(8192, 8192, device=DEV).bfloat16()
    B = torch.randn(8192, 8192, device=DEV).bfloat16()
    sustained(lambda: torch.matmul(A, B), 2.0 * 8192 ** 3, secs, smi, "cublas_bf16_8192^3")
    for K in (4096, 11008):
        for variant in variants:
            for chunk in ((0, 1024) if variant != 1 else (0,)):
                sustained(make_syrk(K, 32768, variant, chunk), issued_flops(K, 32768), secs, smi, f"syrk_v{variant}_K{K}_chunk{chunk or 512}")
    sustained(lambda: torch.matmul(A, B), 2.0 * 8192 ** 3, secs, smi, "cublas_bf16_8192^3_again")
    smi.stop()


if __name__ == "__main__":
    main()
