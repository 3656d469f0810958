#!/usr/bin/env python
"""Benchmark of the GPTQ layer-wise quantization hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): ONE Llama-2-7B
decoder layer -- q,k,v,o 4096x4096, gate,up 11008x4096, down 4096x11008 -- quantized with GPTQ
4-bit, groupsize 128, act-order, static groups, true-sequential subset order, from 128 x 2048
synthetic calibration tokens (random-init fp16 weights N(0, 0.02^2), activations N(0,1) x per-channel
exp(0.5 z)).  A step = the whole hot path for that layer through the reference-facing
Python API: 128 add_batch calls per linear (Hessian), 7 fasterquant calls, 7 packs.
Whole-model time = 32 x step.

`value`  : quantized weights per second, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric with every step's activations and weights starting in PINNED HOST memory
           (H2D copies inside the timed region) and the packed result read back to the host.
N > 1    : strong scaling of the same layer: calibration samples sharded (one NCCL all-reduce of each
           partial Hessian), W rows sharded (one all-gather of weights/codes/scales/zeros).
--impl reference : the reference algorithm on the host CPU (oracle/gptq_oracle.py, a torch-CPU
           restatement pinned bit-exact to the reference), bounded sample per step, extrapolated.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HID, INTER, NSAMPLES, SEQ = 4096, 11008, 128, 2048
# (name, rows, cols, activation-id); activation 0 = attn input, 1 = o_proj input, 2 = mlp input, 3 = down input
LINEARS = [("k_proj", HID, HID, 0), ("v_proj", HID, HID, 0), ("q_proj", HID, HID, 0), ("o_proj", HID, HID, 1),
           ("up_proj", INTER, HID, 2), ("gate_proj", INTER, HID, 2), ("down_proj", HID, INTER, 3)]
SUBSETS = [[0, 1, 2], [3], [4, 5], [6]]                       # llama.py:98-103 true-sequential order
ACT_K = [HID, HID, HID, INTER]
LAYER_WEIGHTS = sum(r * c for _, r, c, _ in LINEARS)
N_LAYERS = 32
GPTQ_KW = dict(blocksize=128, percdamp=.01, groupsize=128, actorder=True, static_groups=True)
WBITS = 4


def syrk_traffic(prof):
    """Launch-weighted DRAM bytes per syrk launch of the timed region, from the committed ncu capture
    (profiles/r01_syrk_traffic.json: per launch shape).  None if a launch shape was not captured."""
    path = os.path.join(ROOT, "profiles", "r01_syrk_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        shapes = json.load(f)["launch_shapes"]
    tot, n = 0.0, 0
    for name, _, _, work, meta in prof:
        if name != "syrk":
            continue
        ent = shapes.get("K%d_T%d" % meta) if meta else None
        if ent is None:
            return None
        tot += ent["dram_read_bytes"] + ent["dram_write_bytes"]
        n += 1
    return tot / n if n else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_sample(threads=None):
    """Bounded sample of the workload on the host CPU with the oracle port; returns extrapolated layer seconds."""
    from oracle import gptq_oracle as O
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    K = HID
    chan = torch.exp(0.5 * torch.randn(K, generator=g))
    x = (torch.randn(1, SEQ, K, generator=g) * chan).half()
    st = O.HessianState(K)
    t0 = time.perf_counter()
    O.add_batch(st, x)                                       # gptq.py:38-65, one of the 128 samples
    t_h = time.perf_counter() - t0
    # fasterquant on a 512-row slice of q_proj (the chain is full size; the loop is row-proportional)
    RS = 512
    W = (torch.randn(RS, K, generator=g) * 0.02).half()
    H = st.H + torch.eye(K) * 1e-3
    t0 = time.perf_counter()
    H2 = H.clone()
    O.hinv_upper(H2, 0.01)
    t_chain = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.fasterquant(W, H, O.QParams(bits=WBITS), **{k: v for k, v in GPTQ_KW.items()})
    t_fq = time.perf_counter() - t0
    t_loop = max(t_fq - t_chain, 1e-6)                       # row-proportional part for RS rows, K=4096
    # extrapolation to the full layer (stated in cpu_baseline.sample):
    #   Hessian: dense fp32 sgemm, time ~ K^2 * tokens  -> 128 samples x (6 x K=4096 + 1 x K=11008)
    #   chain:   ~ K^3                                   -> 6 x K=4096 + 1 x K=11008
    #   loop:    ~ R * K^2                               -> per linear
    kr = (INTER / HID)
    t_hess = NSAMPLES * t_h * (6 + kr ** 2)
    t_ch = t_chain * (6 + kr ** 3)
    t_lp = sum(t_loop * (r / RS) * (c / K) ** 2 for _, r, c, _ in LINEARS)
    total = t_hess + t_ch + t_lp
    return dict(layer_seconds=total, hessian_seconds=t_hess, chain_seconds=t_ch, loop_seconds=t_lp,
                measured=dict(add_batch_1sample_K4096=t_h, chain_K4096=t_chain, fasterquant_512x4096=t_fq),
                cores=torch.get_num_threads(),
                sample=("per step: add_batch on 1 of 128 samples (2048 tokens, K=4096) + fasterquant of a 512x4096 slice of "
                        "q_proj (full 4096^2 inverse-Cholesky chain); extrapolated to the layer with Hessian~K^2*tokens, "
                        "chain~K^3, loop~R*K^2"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ts = []
    last = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_reference_sample()
        if i >= args.warmup:
            ts.append(last["layer_seconds"])
    layer_s = statistics.mean(ts)
    val = LAYER_WEIGHTS / layer_s / 1e6
    out = base_line(args, val, layer_s * 1e3)
    out.update({"impl": "reference", "dtype": "f32", "scaling": "strong",
                "cpu_baseline": {"value": val, "unit": "Mweights/s", "cores": last["cores"], "kind": "port",
                                 "sample": last["sample"], "detail": last["measured"]},
                "e2e": {"value": val, "unit": "Mweights/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "whole_model_seconds": layer_s * N_LAYERS, "gpu_launches": 0})
    print(json.dumps(out))


def base_line(args, value, ms_per_step):
    return {"metric": "llama2_7b_gptq_4b_g128_actorder_layer_quantization_throughput", "value": value, "unit": "Mweights/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16xf16->f32 (Hessian), f32 (solver)",
            "data": "synthetic",
            "config": {"workload": "Llama-2-7B decoder layer (q,k,v,o 4096x4096; gate,up 11008x4096; down 4096x11008), "
                                   "GPTQ 4-bit g128 act-order static-groups true-sequential, 128x2048 calibration tokens; "
                                   "step = 128 add_batch/linear + 7 fasterquant + 7 pack; whole model = 32 steps",
                       "weights_per_step": LAYER_WEIGHTS, "calibration_tokens": NSAMPLES * SEQ,
                       "l2": "inputs (12.2 GB of activations per step) exceed the 126 MB L2; no explicit flush"}}


# ------------------------------------------------------------------------------------------ our arm (GPU)
def make_inputs(dev, rank, world, host=False):
    """Synthetic activations for this rank's sample shard, and the layer's weights."""
    from gptq_b200.dist import sample_shard
    mine = sample_shard(NSAMPLES, rank, world)
    acts = []
    for a, K in enumerate(ACT_K):
        g = torch.Generator(device=dev).manual_seed(100 + a)
        chan = torch.exp(0.5 * torch.randn(K, device=dev, generator=g))
        if host:
            X = torch.empty((len(mine), SEQ, K), dtype=torch.float16, pin_memory=True)
        else:
            X = torch.empty((len(mine), SEQ, K), dtype=torch.float16, device=dev)
        for i, j in enumerate(mine):
            gj = torch.Generator(device=dev).manual_seed(1000 * (a + 1) + j)
            xj = (torch.randn(SEQ, K, device=dev, generator=gj) * chan).half()
            X[i].copy_(xj)
        acts.append(X)
    ws = []
    gw = torch.Generator(device=dev).manual_seed(1)
    for _, r, c, _ in LINEARS:
        w = (torch.randn(r, c, device=dev, generator=gw) * 0.02).half()
        ws.append(w.cpu().pin_memory() if host else w)
    torch.cuda.synchronize()
    return acts, ws


_COPY_STREAMS = {}


def _copy_stream(dev):
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[dev]


_H2D = {}


def _h2d_buffers(dev, nbytes, wbytes):
    """Two device landing areas (one per in-flight subset: activations + that subset's weights) and their 'consumed'
    events, allocated once."""
    st = _H2D.get(dev)
    if st is None or st["bufs"][0].numel() < nbytes or st["wbufs"][0].numel() < wbytes:
        st = {"bufs": [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)],
              "wbufs": [torch.empty(wbytes, dtype=torch.uint8, device=dev) for _ in range(2)],
              "consumed": [None, None]}
        _H2D[dev] = st
        # fresh blocks of the compute stream's allocator pool: the copy stream may only touch them after everything
        # already queued on the compute stream (possible previous users of that memory) has run
        _copy_stream(dev).wait_stream(torch.cuda.current_stream())
    return st


def _enqueue_subset_h2d(dev, X, wlist, slot):
    """Caller side of the public API when activations and weights live in pinned host memory: stream the samples of
    one subset, then its weights, into landing area `slot` on the copy stream (one event per sample, one for the
    weights).  Everything host-to-device goes through this one stream: a copy issued on the compute stream would
    queue behind the transfers already in flight and stall the kernels behind it."""
    cs = _copy_stream(dev)
    wmax = max(sum(LINEARS[li][1] * LINEARS[li][2] for li in sub) for sub in SUBSETS) * 2
    st = _h2d_buffers(dev, NSAMPLES * SEQ * max(ACT_K) * 2, wmax)
    n = X.numel() * X.element_size()
    view = st["bufs"][slot][:n].view(X.dtype).view(X.shape)
    ready, wdev = [], []
    with torch.cuda.stream(cs):
        if st["consumed"][slot] is not None:
            cs.wait_event(st["consumed"][slot])           # the previous occupant of this area is done with it
        for j in range(X.shape[0]):
            view[j].copy_(X[j], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            ready.append(ev)
        off = 0
        for w in wlist:
            nb = w.numel() * w.element_size()
            wd = st["wbufs"][slot][off:off + nb].view(w.dtype).view(w.shape)
            wd.copy_(w, non_blocking=True)
            wdev.append(wd)
            off += (nb + 255) // 256 * 256
        wev = torch.cuda.Event()
        wev.record(cs)
    return view, ready, wdev, wev


def quantize_layer(dev, acts, weights, world, from_host, carry=None, prefetch_next=False):
    """One step: the reference driver's per-layer flow (llama.py:97-170) over our GPTQ / Quantizer / pack."""
    import torch.nn as nn
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer, pack_codes
    from gptq_b200 import dist as gd
    results = []
    d2h = 0
    cur = torch.cuda.current_stream()
    pending = None
    trace = [] if (from_host and os.environ.get("GPTQ_B200_E2E_TRACE")) else None

    def mark(name, stream):
        if trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            trace.append((name, ev))
    mark("step start", cur)
    if from_host:
        # `carry`: the previous step already queued this step's first subset behind its own last transfer
        pending = carry if carry is not None else _enqueue_subset_h2d(
            dev, acts[LINEARS[SUBSETS[0][0]][3]], [weights[li] for li in SUBSETS[0]], 0)
    for si, subset in enumerate(SUBSETS):
        gs = []
        for li in subset:
            name, r, c, a = LINEARS[li]
            lin = nn.Linear(c, r, bias=False, device="meta")
            lin.to_empty(device=dev)
            lin = lin.half()
            if not from_host:
                lin.weight.data = weights[li].clone()
            g = GPTQ(lin)
            g.quantizer = Quantizer()
            g.quantizer.configure(WBITS, perchannel=True, sym=False, mse=False)
            gs.append(g)
        X = acts[LINEARS[subset[0]][3]]
        if from_host:
            # the H2D stream runs one subset ahead of the compute stream (while this subset's Hessians, chain and
            # column loops run, the next subset's samples are already landing in the other buffer)
            Xd, ready, wdev, wev = pending
            if si == 0:
                mark("h2d subset 0 done", _copy_stream(dev))
            if si + 1 < len(SUBSETS) or prefetch_next:      # the upload keeps running one subset ahead, across steps too
                nxt = SUBSETS[(si + 1) % len(SUBSETS)]
                pending = _enqueue_subset_h2d(dev, acts[LINEARS[nxt[0]][3]], [weights[li] for li in nxt], (si + 1) & 1)
                mark(f"h2d subset {(si + 1) % len(SUBSETS)} done", _copy_stream(dev))
            else:
                pending = None
            for j in range(Xd.shape[0]):
                cur.wait_event(ready[j])
                xj = Xd[j].unsqueeze(0)
                for g in gs:
                    g.add_batch(xj, None)
            mark(f"hessian subset {si} done", cur)
            cur.wait_event(wev)                           # this subset's weights have landed
            for g, wd in zip(gs, wdev):
                g.layer.weight.data = wd
        else:
            for j in range(X.shape[0]):
                xj = X[j].unsqueeze(0)
                for g in gs:                              # the forward hooks of q/k/v see the same tensor
                    g.add_batch(xj, None)
        if os.environ.get("GPTQ_B200_DEBUG_SYNC"):
            torch.cuda.synchronize()
            fin = [bool(torch.isfinite(g.raw_accumulator()).all()) for g in gs]
            print("debug: add_batch phase ok", [LINEARS[li][0] for li in subset], "S finite:", fin, file=sys.stderr)
        for g in gs:
            if world > 1:
                gd.allreduce_hessian(g)
                gd.fasterquant_sharded(g, **GPTQ_KW)
            else:
                g.fasterquant(**GPTQ_KW)
            qw = pack_codes(g.codes, WBITS)
            if from_host:                                 # result back to the host: packed words + scales + zeros
                outs = [qw.cpu(), g.scales.cpu(), g.zeros.cpu(), g.g_idx.cpu()]
                d2h += sum(o.numel() * o.element_size() for o in outs)
            results.append((qw, g.error))
            g.free()
        if from_host:                                     # activations staged, weights read: the landing area is free
            ev = torch.cuda.Event()
            ev.record(cur)
            _H2D[dev]["consumed"][si & 1] = ev
        mark(f"fasterquant subset {si} done", cur)
    if from_host and pending is None:
        cur.wait_stream(_copy_stream(dev))
    if trace is not None:
        torch.cuda.synchronize()
        t0 = trace[0][1]
        print("e2e trace (ms since step start): " + "; ".join(f"{n} {t0.elapsed_time(e):.1f}" for n, e in trace[1:]), file=sys.stderr)
    return results, d2h, pending


def run_ours(args):
    import contextlib
    import io
    import torch.distributed as dist
    from gptq_b200 import _lib
    import gptq_b200.gptq as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(acts, weights, from_host, steps, warmup, profile):
        quiet = io.StringIO()
        d2h = 0
        with contextlib.redirect_stdout(quiet):
            for _ in range(warmup):
                quantize_layer(dev, acts, weights, world, from_host)
            carry = None
            barrier()
            G.PROFILE = [] if profile else None
            n0 = _lib.launch_count()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for i in range(steps):
                _, d2h, carry = quantize_layer(dev, acts, weights, world, from_host, carry, prefetch_next=from_host and i + 1 < steps)
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        prof = G.PROFILE
        G.PROFILE = None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, _lib.launch_count() - n0, prof, d2h

    # ---- device-resident run (value)
    acts, weights = make_inputs(dev, rank, world, host=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall, launches, prof, _ = timed(acts, weights, False, args.steps, args.warmup, True)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    value = LAYER_WEIGHTS / (ms_step / 1e3) / 1e6

    # per-kernel device time inside the timed region (CUDA events on the launching stream)
    agg = {}
    for name, a, b, work, _meta in prof:
        d = agg.setdefault(name, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += work
        d[2] += 1
    sy = agg.get("syrk", [1e-9, 0.0, 1])
    syrk_tflops = sy[1] / (sy[0] / 1e3) / 1e12
    dense_equiv = sum(2.0 * c * c * (NSAMPLES // world) * SEQ for _, r, c, _ in LINEARS) * args.steps / (sy[0] / 1e3) / 1e12
    roofline = {"kernel": "syrk_tc_kernel (tcgen05 Hessian accumulation)", "bound": "tensor", "achieved": syrk_tflops,
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": syrk_tflops / pk["tf_sustained"],
                "peak_source": pk["source"] + " bf16 cuBLAS sustained", "frac_of_nominal_2250": syrk_tflops / 2250.0,
                "launches": sy[2], "avg_launch_ms": sy[0] / max(sy[2], 1), "share_of_step": sy[0] / ms,
                "dense_equivalent_tflops": dense_equiv, "traffic": syrk_traffic(prof) if world == 1 else None,
                "traffic_note": "DRAM read+write bytes per launch (launch-weighted over the K=4096 and K=11008 launches) from the "
                                "committed ncu capture profiles/r01_syrk_traffic.json; algorithmic bytes per launch: 0.40 GB / 1.22 GB"}
    co = agg.get("colloop", [1e-9, 0.0, 1])
    hv = agg.get("hinv_upper", [1e-9, 0.0, 1])
    others = {
        "colloop": {"bound": "hbm", "achieved": co[1] / (co[0] / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": co[1] / (co[0] / 1e3) / 1e9 / pk["hbm_gbs"], "share_of_step": co[0] / ms,
                    "us_per_column": co[0] * 1e3 / (args.steps * sum(c for _, r, c, _ in LINEARS)),
                    "note": "sequential column recurrence: latency-bound, low HBM fraction expected (SURVEY 8d)"},
        "hinv_upper": {"bound": "fp32", "achieved": hv[1] / (hv[0] / 1e3) / 1e12, "unit": "TFLOP/s (2/3 K^3 useful)",
                       "share_of_step": hv[0] / ms},
    }
    phase_ms = {k: v[0] / args.steps for k, v in agg.items()}
    phase_ms["unattributed (staging copies, pack, allocator, host gaps)"] = ms_step - sum(phase_ms.values())
    del acts, weights
    torch.cuda.empty_cache()

    # ---- end-to-end run: host buffers, H2D/D2H inside the timed region
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        acts_h, weights_h = make_inputs(dev, rank, world, host=True)
        h2d = sum(x.numel() * 2 for x in acts_h) + sum(w.numel() * 2 for w in weights_h)
        e_steps = max(1, min(args.steps, 3))
        ems, _, _, _, d2h = timed(acts_h, weights_h, True, e_steps, 1, False)
        e2e = {"value": LAYER_WEIGHTS / (ems / e_steps / 1e3) / 1e6, "unit": "Mweights/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": ems / e_steps, "steps": e_steps}
        del acts_h, weights_h
    except Exception as ex:  # pinned allocation can fail on small hosts; report instead of hiding
        if not args.no_e2e:
            import traceback
            traceback.print_exc(file=sys.stderr)
        e2e = {"value": None, "unit": "Mweights/s", "error": repr(ex)}

    if rank == 0:
        out = base_line(args, value, ms_step)
        out.update({"whole_model_seconds": ms_step * N_LAYERS / 1e3, "wall_seconds_timed_region": wall,
                    "hessian_tflops_issued": syrk_tflops, "hessian_tflops_dense_equivalent": dense_equiv,
                    "roofline": roofline, "rooflines_other": others, "phase_ms_per_step": phase_ms, "clocks": clocks, "e2e": e2e,
                    "gpu_launches": launches, "impl": "ours"})
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample()
            out["cpu_baseline"] = {"value": LAYER_WEIGHTS / cb["layer_seconds"] / 1e6, "unit": "Mweights/s", "cores": cb["cores"],
                                   "kind": "port", "sample": cb["sample"], "detail": cb["measured"],
                                   "extrapolated_layer_seconds": cb["layer_seconds"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) != args.gpus and args.gpus > 1:
            # convenience: spawn torchrun ourselves when launched as plain `python bench.py --gpus N`
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
