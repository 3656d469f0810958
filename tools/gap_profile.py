"""Device timeline of ONE bench step (torch.profiler / CUPTI): where the GPU idles and which kernels fill the step.

    python tools/gap_profile.py [config] > gpurun_out/gaps.json

Prints: busy / idle totals, device time per kernel name, and the largest idle gaps with the kernels on either side.
(CUPTI tracing slows the host down; shares and gap LOCATIONS are what to read, not absolute milliseconds.)
"""
import contextlib
import io
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "llama7b"
    wl = bench.WORKLOADS[cfg]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    acts, weights = bench.make_inputs(wl, dev, 0, 1, host=False)
    quiet = io.StringIO()
    with contextlib.redirect_stdout(quiet):
        for _ in range(2):
            bench.quantize_layer(wl, dev, acts, weights, 1, False)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        with contextlib.redirect_stdout(quiet):
            bench.quantize_layer(wl, dev, acts, weights, 1, False)
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            evs.append((e.time_range.start, e.time_range.end, e.name))
    evs.sort()
    if not evs:
        print(json.dumps({"error": "no device events"}))
        return
    t0, t1 = evs[0][0], max(e[1] for e in evs)
    busy_end = evs[0][0]
    busy = 0.0
    gaps = []
    per = {}
    prev = None
    for s, e, n in evs:
        short = n.split("(")[0][-60:]
        d = per.setdefault(short, [0.0, 0])
        d[0] += e - s
        d[1] += 1
        if s > busy_end:
            gaps.append((s - busy_end, prev, short, busy_end - t0))
            busy += 0.0
        busy += max(0.0, e - max(s, busy_end))
        if e > busy_end:
            busy_end = e
            prev = short
    gaps.sort(reverse=True)
    total = t1 - t0
    hist = {"<5us": 0.0, "5-20us": 0.0, "20-100us": 0.0, "100us-1ms": 0.0, ">1ms": 0.0}
    for g, *_ in gaps:
        k = "<5us" if g < 5 else "5-20us" if g < 20 else "20-100us" if g < 100 else "100us-1ms" if g < 1000 else ">1ms"
        hist[k] += g
    # idle time attributed to the kernel that FOLLOWS the gap (who was late)
    late = {}
    for g, p, n, _ in gaps:
        late[n] = late.get(n, 0.0) + g
    out = {"config": cfg, "step_ms": total / 1e3, "busy_ms": busy / 1e3, "idle_ms": (total - busy) / 1e3, "n_kernels": len(evs),
           "idle_by_gap_size_ms": {k: v / 1e3 for k, v in hist.items()},
           "idle_before_kernel_ms": {k: round(v / 1e3, 3) for k, v in sorted(late.items(), key=lambda kv: -kv[1])[:20]},
           "device_ms_by_kernel": {k: [round(v[0] / 1e3, 3), v[1]] for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:40]},
           "largest_gaps": [{"gap_us": round(g, 1), "after": p, "before": n, "at_ms": round(at / 1e3, 2)} for g, p, n, at in gaps[:30]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
