"""Where do occasional long steps of the layer bench come from?  Runs bench.quantize_layer in a loop with host-side timers
around the main calls and prints every call that took unusually long on the host, with the device time of its step.

    python tools/stall_probe.py [steps]
"""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import gptq_b200.gptq as G  # noqa: E402
from gptq_b200 import _lib  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if os.environ.get("GPTQ_B200_EXPANDABLE_SEGMENTS", "1") == "1":
    torch.cuda.memory._set_allocator_settings("expandable_segments:True")
lib = _lib.load()
wl = bench.WORKLOADS["llama7b"]
acts, weights = bench.make_inputs(wl, dev, 0, 1, host=False)
LOG = []
STEP = [0]


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            dt = (time.perf_counter() - t0) * 1e3
            LOG.append((STEP[0], label, dt))
    setattr(obj, name, w)


wrap(G, "_solve_hessian", "solve_hessian(host)")
wrap(G, "_solve_rows", "solve_rows(host)")
wrap(G.GPTQ, "fasterquant", "fasterquant(host, includes the record read-back)")
wrap(G.GPTQ, "add_batch", "add_batch(host)")
import gc
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(4):
        bench.quantize_layer(wl, dev, acts, weights, 1, False)
gc.collect(); gc.freeze(); gc.disable()
torch.cuda.synchronize()
LOG.clear()
marks = []
e0 = torch.cuda.Event(enable_timing=True); e0.record()
host = []
with contextlib.redirect_stdout(io.StringIO()):
    for i in range(steps):
        STEP[0] = i
        t0 = time.perf_counter()
        bench.quantize_layer(wl, dev, acts, weights, 1, False)
        host.append((time.perf_counter() - t0) * 1e3)
        ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append(ev)
torch.cuda.synchronize()
dev_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(steps)]
print("device ms per step:", [round(x, 1) for x in dev_ms])
print("host   ms per step:", [round(x, 1) for x in host])
# typical duration per label
import statistics
by = {}
for s, label, dt in LOG:
    by.setdefault(label, []).append(dt)
for label, v in by.items():
    print(f"{label}: n={len(v)} median={statistics.median(v):.2f} ms max={max(v):.2f} ms")
for s, label, dt in LOG:
    med = statistics.median(by[label])
    if dt > 5 * med + 5:
        print(f"  OUTLIER step {s}: {label} took {dt:.1f} ms (median {med:.2f})")
