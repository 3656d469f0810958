#!/usr/bin/env python
"""Benchmark of the GPTQ layer-wise quantization hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cuda]
                    [--config llama7b|opt125m|opt1.3b|bloom7b1|opt66b-layer] [--workload layer|model]

Default workload (BASELINE.json configs[2], the configuration the metric is quoted on): ONE Llama-2-7B decoder layer --
q,k,v,o 4096x4096, gate,up 11008x4096, down 4096x11008 -- quantized with GPTQ 4-bit, groupsize 128, act-order, static
groups, true-sequential subset order, from 128 x 2048 synthetic calibration tokens (random-init fp16 weights
N(0, 0.02^2), activations N(0,1) x per-channel exp(0.5 z)).  A step = the whole hot path for that layer through
the reference-facing Python API: 128 add_batch calls per linear (Hessian), one fasterquant and one pack per linear.
`--config` selects the decoder layer of another BASELINE config (same definition of a step).

`value`  : quantized weights per second, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric with every step's activations and weights starting in PINNED HOST memory (H2D copies inside the
           timed region) and the packed result read back to the host.
N > 1    : strong scaling of the same layer: calibration samples sharded (one all-reduce of each partial Hessian's lower
           triangle), W rows sharded (one coalesced all-gather of weights/codes/scales/zeros); a `parity` object compares
           the sharded result of one linear with the single-GPU result (outside the timed region).
--impl reference      : the UNMODIFIED reference (baseline/_ref, see baseline/install_reference.py) on the host CPU with
           all host threads; each step is a bounded sample of the layer; the layer time is assembled from the measured
           unit costs (stated in `cpu_baseline.sample`).  Falls back to the oracle port if baseline/_ref is absent.
--impl reference-cuda : the same unmodified reference classes with CUDA tensors on the same GPU (cuBLAS fp32 sgemm,
           cuSOLVER, eager column loop; TF32 off) -- the like-for-like GPU baseline, whole layer, no sampling.
--workload model      : whole-model quantization through gptq_b200.sequential (random-init HF model; SURVEY 8d definition
           of "whole-model quantize time"), see run_model().
"""
from __future__ import annotations

import argparse
import contextlib
import gc
import io
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

NSAMPLES, SEQ = 128, 2048


def _layer(title, linears, subsets, act_k, n_layers, wbits, gptq, pack="codes", note=""):
    return dict(title=title, linears=linears, subsets=subsets, act_k=act_k, n_layers=n_layers, wbits=wbits, gptq=gptq,
                pack=pack, note=note, weights=sum(r * c for _, r, c, _ in linears))


def _opt(h, f):
    return [("k_proj", h, h, 0), ("v_proj", h, h, 0), ("q_proj", h, h, 0), ("out_proj", h, h, 1), ("fc1", f, h, 2), ("fc2", h, f, 3)]


# (name, rows, cols, activation-id): linears with the same activation-id are hooked on the same tensor (share one Hessian)
WORKLOADS = {
    # configs[2] -- llama.py:98-103 true-sequential subset order
    "llama7b": _layer("Llama-2-7B decoder layer (q,k,v,o 4096x4096; gate,up 11008x4096; down 4096x11008), GPTQ 4-bit g128 "
                      "act-order static-groups true-sequential",
                      [("k_proj", 4096, 4096, 0), ("v_proj", 4096, 4096, 0), ("q_proj", 4096, 4096, 0), ("o_proj", 4096, 4096, 1),
                       ("up_proj", 11008, 4096, 2), ("gate_proj", 11008, 4096, 2), ("down_proj", 4096, 11008, 3)],
                      [[0, 1, 2], [3], [4, 5], [6]], [4096, 4096, 4096, 11008], 32, 4,
                      dict(blocksize=128, percdamp=.01, groupsize=128, actorder=True, static_groups=True)),
    # configs[0] -- opt.py:109-187: all six linears of a layer are hooked in ONE forward (no true-sequential)
    "opt125m": _layer("OPT-125m decoder layer (q,k,v,out 768x768; fc1 3072x768; fc2 768x3072), GPTQ 4-bit per-channel",
                      _opt(768, 3072), [[0, 1, 2, 3, 4, 5]], [768, 768, 768, 3072], 12, 4,
                      dict(blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=True),
                      note="static_groups=True is opt.py's default (opt.py:584-587), a no-op at groupsize -1"),
    # configs[1] -- packed through the fp16-weights entry point Quant3Linear.pack(linear, scales, zeros) (quant.py:152-187)
    "opt1.3b": _layer("OPT-1.3b decoder layer (q,k,v,out 2048x2048; fc1 8192x2048; fc2 2048x8192), GPTQ 3-bit g128 + "
                      "Quant3Linear.pack", _opt(2048, 8192), [[0, 1, 2, 3, 4, 5]], [2048, 2048, 2048, 8192], 24, 3,
                      dict(blocksize=128, percdamp=.01, groupsize=128, actorder=False, static_groups=True), pack="linear",
                      note="static_groups=True is opt.py's default (opt.py:584-587)"),
    # configs[3] -- bloom.py:93-170, dynamic groups (stale-W group fit, gptq.py:253-255)
    "bloom7b1": _layer("BLOOM-7b1 decoder layer (query_key_value 12288x4096; dense 4096x4096; dense_h_to_4h 16384x4096; "
                       "dense_4h_to_h 4096x16384), GPTQ 2-bit g64",
                       [("query_key_value", 12288, 4096, 0), ("dense", 4096, 4096, 1), ("dense_h_to_4h", 16384, 4096, 2),
                        ("dense_4h_to_h", 4096, 16384, 3)], [[0, 1, 2, 3]], [4096, 4096, 4096, 16384], 30, 2,
                       dict(blocksize=128, percdamp=.01, groupsize=64, actorder=False, static_groups=False)),
    # configs[4] -- one decoder layer resident at a time
    "opt66b-layer": _layer("OPT-66B decoder layer (q,k,v,out 9216x9216; fc1 36864x9216; fc2 9216x36864), GPTQ 3-bit act-order",
                           _opt(9216, 36864), [[0, 1, 2, 3, 4, 5]], [9216, 9216, 9216, 36864], 64, 3,
                           dict(blocksize=128, percdamp=.01, groupsize=-1, actorder=True, static_groups=True),
                           note="static_groups=True is opt.py's default, a no-op at groupsize -1"),
}


def syrk_traffic(prof):
    """Launch-weighted DRAM bytes per Hessian-kernel launch of the timed region, from the committed ncu capture
    (profiles/*_syrk_traffic.json: per launch shape).  None if a launch shape was not captured."""
    path = None
    for name in ("r02_syrk_traffic.json", "r01_syrk_traffic.json"):
        if os.path.exists(os.path.join(ROOT, "profiles", name)):
            path = os.path.join(ROOT, "profiles", name)
            break
    if path is None:
        return None, None
    with open(path) as f:
        shapes = json.load(f)["launch_shapes"]
    tot, n = 0.0, 0
    for name, _, _, work, meta in prof:
        if name != "syrk":
            continue
        ent = shapes.get("K%d_T%d" % meta) if meta else None
        if ent is None:
            return None, os.path.basename(path)
        tot += ent["dram_read_bytes"] + ent["dram_write_bytes"]
        n += 1
    return (tot / n if n else None), os.path.basename(path)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def numa_bind(local):
    """Multi-GPU runs: pin this rank's main thread to the CPUs of the NUMA node its GPU hangs off BEFORE the pinned host
    buffers are allocated, so that their pages are local to that node (first touch) and the rank's host-to-device stream does
    not cross the socket interconnect.  What a multi-GPU calibration driver does with its pinned staging memory; a no-op when
    the platform exposes no NUMA information (numa_node = -1), when the node has no CPU this process may use, or with
    GPTQ_B200_NUMA_BIND=0.  Returns a description for the bench line."""
    info = {"applied": False}
    if os.environ.get("GPTQ_B200_NUMA_BIND", "1") != "1":
        info["note"] = "disabled (GPTQ_B200_NUMA_BIND=0)"
        return info
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        info["gpu_pci"] = bdf
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & allowed)
        info["node_cpus_allowed"] = len(use)
        info["cpus_allowed_before"] = len(allowed)
        if not use:
            return info
        os.sched_setaffinity(0, use)
        info["applied"] = True
    except Exception as ex:
        info["error"] = repr(ex)
    return info


class ClockSampler:
    """SM clock and clock-event reasons of GPU `index`, sampled DURING the timed region.
    GPTQ_B200_CLOCK_SAMPLER = nvml (default: NVML queries from a thread of this process every 100 ms), smi (an `nvidia-smi
    -lms 200` child process, the round-1 sampler) or off."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.kind = os.environ.get("GPTQ_B200_CLOCK_SAMPLER", "nvml")
        self.thread, self._stop, self.error = None, threading.Event(), None

    def start(self):
        if self.kind == "off":
            return
        if self.kind == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nv = pynvml
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
                self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
                return
            except Exception as ex:              # fall back to the child process
                self.error = repr(ex)
                self.kind = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                t0 = time.perf_counter()
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                t1 = time.perf_counter()
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                t2 = time.perf_counter()
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                t3 = time.perf_counter()
                self.call_ms = [max(a, b) for a, b in zip(getattr(self, "call_ms", [0, 0, 0]), [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3])]
                self.rows.append([str(sm), str(self.mx), str(pw)] + ["Active" if (r & b) else "Not Active" for b in bits])
            except Exception as ex:
                self.error = repr(ex)
            self._stop.wait(0.1)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.kind == "off":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off (GPTQ_B200_CLOCK_SAMPLER=off)"]}
        if self.kind == "nvml":
            self._stop.set()
            if self.thread is not None:
                self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "source": "NVML, in-process thread, 100 ms" if self.kind == "nvml" else "nvidia-smi -lms 200",
                "nvml_call_ms_max[clock,reasons,power]": [round(x, 2) for x in getattr(self, "call_ms", [])]}


def metric_name(cfg):
    if cfg == "llama7b":
        return "llama2_7b_gptq_4b_g128_actorder_layer_quantization_throughput"
    return cfg.replace("-", "_").replace(".", "_") + "_gptq_layer_quantization_throughput"


def base_line(args, wl, value, ms_per_step):
    return {"metric": metric_name(args.config), "value": value, "unit": "Mweights/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16xf16->f32 (Hessian), f32 (solver)",
            "data": "synthetic",
            "config": {"workload": wl["title"] + f", {NSAMPLES}x{SEQ} calibration tokens; step = {NSAMPLES} add_batch/linear + "
                                   f"{len(wl['linears'])} fasterquant + {len(wl['linears'])} pack; whole model = {wl['n_layers']} such layers",
                       "weights_per_step": wl["weights"], "calibration_tokens": NSAMPLES * SEQ, "note": wl["note"],
                       "l2": "inputs (GBs of activations per step) exceed the 126 MB L2; no explicit flush",
                       "gc": "Python's cyclic GC is disabled inside the timed region (gc.freeze + gc.disable), nothing else differs "
                             "from a plain driver loop"}}


# ------------------------------------------------------------------------------------------ reference arms
def _reference_classes():
    """(GPTQ class, Quantizer class, kind): the unmodified reference from baseline/_ref, else the oracle port."""
    from baseline import reference as BR
    if BR.available():
        rg, rq = BR.load()
        return rg.GPTQ, rq.Quantizer, "reference"
    from oracle.ref_gptq import OracleGPTQ, OracleQuantizer
    return OracleGPTQ, OracleQuantizer, "port"


def _host_threads():
    """All host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _ref_gptq(GPTQcls, Qcls, W, wbits, dev="cpu"):
    import torch.nn as nn
    R, K = W.shape
    lin = nn.Linear(K, R, bias=False)
    lin.weight.data = W.to(dev)
    lin = lin.to(dev)
    g = GPTQcls(lin)
    g.quantizer = Qcls()
    g.quantizer.configure(wbits, perchannel=True, sym=False, mse=False)
    return lin, g


_CPU_ONCE = {}


def cpu_reference_sample(wl):
    """One bounded sample of the layer on the host CPU with the reference's own classes.  Measured every call: add_batch
    on 2 of the 128 samples and one FULL fasterquant at the layer's first linear shape.  Measured once per process
    (seconds each, reused by later steps): a 512-row fasterquant at that shape (separates the row-proportional column loop
    from the Cholesky chain), and for every other distinct K of the layer one add_batch sample and fasterquant of two row
    slices.  The layer time is assembled from those unit costs with the layer's own counts -- no exponent laws."""
    GPTQcls, Qcls, kind = _reference_classes()
    cores = _host_threads()
    quiet = io.StringIO()
    g0 = torch.Generator().manual_seed(0)
    kw = {k: wl["gptq"][k] for k in ("blocksize", "percdamp", "groupsize", "actorder", "static_groups")}
    t_wall0 = time.perf_counter()

    def acts(K, n):
        chan = torch.exp(0.5 * torch.randn(K, generator=g0))
        return [(torch.randn(1, SEQ, K, generator=g0) * chan).half() for _ in range(n)]

    def time_fq(R, K, xs):
        W = (torch.randn(R, K, generator=g0) * 0.02).half()
        _, g = _ref_gptq(GPTQcls, Qcls, W, wl["wbits"])
        th = []
        for x in xs:
            t0 = time.perf_counter()
            g.add_batch(x, None)
            th.append(time.perf_counter() - t0)
        H = g.H if not callable(getattr(g, "H", None)) else None
        if H is not None and hasattr(H, "diagonal"):
            H.diagonal().add_(1e-2 * float(H.diagonal().mean()))     # few samples: keep the sample's H safely PD
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(quiet), contextlib.redirect_stderr(quiet):
            g.fasterquant(**kw)
        return statistics.mean(th), time.perf_counter() - t0

    Ks = sorted({c for _, r, c, _ in wl["linears"]})
    K0 = wl["linears"][0][2]
    R0 = wl["linears"][0][1]
    key = id(wl)
    once = _CPU_ONCE.setdefault(key, {})
    # every call: the first linear in full
    x0 = acts(K0, 2)
    t_h0, t_full0 = time_fq(R0, K0, x0)
    if "slices" not in once:
        sl = {}
        rs0 = max(128, R0 // 8)
        _, t_s = time_fq(rs0, K0, x0[:1])
        b = max((t_full0 - t_s) / (R0 - rs0), 0.0)
        sl[K0] = dict(a=max(t_full0 - b * R0, 0.0), b=b, t_h=None, measured={f"fasterquant_{rs0}x{K0}": t_s})
        for K in Ks:
            if K == K0:
                continue
            xs = acts(K, 1)
            ra, rb = 128, 640
            th, ta = time_fq(ra, K, xs)
            _, tb = time_fq(rb, K, xs)
            b = max((tb - ta) / (rb - ra), 0.0)
            sl[K] = dict(a=max(ta - b * ra, 0.0), b=b, t_h=th, measured={f"add_batch_1sample_K{K}": th, f"fasterquant_{ra}x{K}": ta,
                                                                         f"fasterquant_{rb}x{K}": tb})
        once["slices"] = sl
    sl = once["slices"]
    # refresh the per-call part of the model for K0
    b0 = sl[K0]["b"]
    a0 = max(t_full0 - b0 * R0, 0.0)
    t_hess = t_solve = 0.0
    for _, r, c, _ in wl["linears"]:
        th = t_h0 if c == K0 else sl[c]["t_h"]
        a, b = (a0, b0) if c == K0 else (sl[c]["a"], sl[c]["b"])
        t_hess += NSAMPLES * th
        t_solve += a + b * r
    measured = {f"add_batch_1sample_K{K0}": t_h0, f"fasterquant_{R0}x{K0}_full": t_full0}
    for K in sl:
        measured.update(sl[K]["measured"])
    return dict(layer_seconds=t_hess + t_solve, hessian_seconds=t_hess, solve_seconds=t_solve, measured=measured, cores=cores,
                kind=kind, sample_wall_seconds=time.perf_counter() - t_wall0,
                sample=(f"per step: reference add_batch on 2 of {NSAMPLES} samples ({SEQ} tokens, K={K0}) + one FULL reference fasterquant "
                        f"of {wl['linears'][0][0]} {R0}x{K0}; once per process: fasterquant of a {max(128, R0 // 8)}-row slice at K={K0} and, "
                        f"for each other K of the layer, one add_batch sample + fasterquant of 128- and 640-row slices (chain in full); layer "
                        f"time = sum over the layer's linears of {NSAMPLES} x add_batch(K) + chain(K) + rows x per-row loop cost(K)"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config]
    ts, walls, last = [], [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(wl)
        if i >= args.warmup:
            ts.append(last["layer_seconds"])
            walls.append(last["sample_wall_seconds"])
    layer_s = statistics.mean(ts)
    val = wl["weights"] / layer_s / 1e6
    out = base_line(args, wl, val, statistics.mean(walls) * 1e3)
    out.update({"impl": "reference", "dtype": "f32", "scaling": "strong",
                "ms_per_step_note": "wall time of one bounded SAMPLE step; the layer itself takes extrapolated_layer_seconds",
                "extrapolated_layer_seconds": layer_s,
                "cpu_baseline": {"value": val, "unit": "Mweights/s", "cores": last["cores"], "kind": last["kind"],
                                 "sample": last["sample"], "detail": last["measured"],
                                 "hessian_seconds": last["hessian_seconds"], "solve_seconds": last["solve_seconds"]},
                "e2e": {"value": val, "unit": "Mweights/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "hot_path_seconds_x_layers": layer_s * wl["n_layers"], "gpu_launches": 0})
    print(json.dumps(out))


def reference_cuda_layer(wl, dev, steps, warmup):
    """The unmodified reference classes with CUDA tensors: whole layer, every sample, every linear (no sharing, no sampling)."""
    GPTQcls, Qcls, kind = _reference_classes()
    acts, weights = make_inputs(wl, dev, 0, 1, host=False)
    kw = {k: wl["gptq"][k] for k in ("blocksize", "percdamp", "groupsize", "actorder", "static_groups")}
    quiet = io.StringIO()
    times, hess, solve = [], [], []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        t_h = t_s = 0.0
        for subset in wl["subsets"]:
            gs = []
            for li in subset:
                _, g = _ref_gptq(GPTQcls, Qcls, weights[li].clone(), wl["wbits"], dev)
                gs.append(g)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
            a.record()
            for j in range(NSAMPLES):
                for li, g in zip(subset, gs):
                    g.add_batch(acts[wl["linears"][li][3]][j].unsqueeze(0), None)
            b.record()
            with contextlib.redirect_stdout(quiet), contextlib.redirect_stderr(quiet):
                for g in gs:
                    g.fasterquant(**kw)
                    g.free()
            c.record()
            torch.cuda.synchronize()
            t_h += a.elapsed_time(b)
            t_s += b.elapsed_time(c)
        e[1].record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e[0].elapsed_time(e[1])); hess.append(t_h); solve.append(t_s)
    del acts, weights
    torch.cuda.empty_cache()
    ms = statistics.mean(times)
    return dict(value=wl["weights"] / (ms / 1e3) / 1e6, unit="Mweights/s", ms_per_step=ms, hessian_ms=statistics.mean(hess),
                solve_ms=statistics.mean(solve), steps=steps, warmup=warmup, kind=kind + " classes, CUDA tensors (cuBLAS sgemm / cuSOLVER / "
                "eager column loop, TF32 off)", note="whole layer, all 128 samples, no sampling or extrapolation")


def run_reference_cuda(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config]
    torch.cuda.set_device(0)
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = reference_cuda_layer(wl, torch.device("cuda", 0), steps, warmup)
    out = base_line(args, wl, r["value"], r["ms_per_step"])
    out.update({"impl": "reference-cuda", "dtype": "f32", "steps": steps, "warmup": warmup, "n_gpus": 1, "gpu_torch_baseline": r,
                "e2e": {"value": None, "unit": "Mweights/s", "note": "device-resident only"}, "gpu_launches": 0,
                "hot_path_seconds_x_layers": r["ms_per_step"] * wl["n_layers"] / 1e3})
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ our arm (GPU)
def make_inputs(wl, dev, rank, world, host=False, all_samples=False):
    """Synthetic activations for this rank's sample shard, and the layer's weights."""
    from gptq_b200.dist import sample_shard
    mine = list(range(NSAMPLES)) if all_samples else sample_shard(NSAMPLES, rank, world)
    acts = []
    for a, K in enumerate(wl["act_k"]):
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
    for _, r, c, _ in wl["linears"]:
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
_D2H = {}
_LAYERS = {}


def _d2h(dev, key, t):
    """Queue a device-to-host copy of a result into a pinned buffer (allocated once per result); returns its bytes."""
    buf = _D2H.get((dev, key))
    if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
        buf = _D2H[(dev, key)] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    return t.numel() * t.element_size()


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


def _subset_acts(wl, subset):
    """Distinct activation ids of a subset, in first-use order."""
    ids = []
    for li in subset:
        a = wl["linears"][li][3]
        if a not in ids:
            ids.append(a)
    return ids


def _enqueue_subset_h2d(wl, dev, acts, weights, subset, slot, wrows=None):
    """Caller side of the public API when activations and weights live in pinned host memory: stream the samples of
    one subset, then its weights, into landing area `slot` on the copy stream (one event per sample, one for the
    weights).  Everything host-to-device goes through this one stream: a copy issued on the compute stream would
    queue behind the transfers already in flight and stall the kernels behind it."""
    cs = _copy_stream(dev)
    ids = _subset_acts(wl, subset)
    n_local = acts[0].shape[0]
    amax = max(sum(n_local * SEQ * wl["act_k"][a] * 2 for a in _subset_acts(wl, sub)) for sub in wl["subsets"])
    wmax = max(sum(wl["linears"][li][1] * wl["linears"][li][2] for li in sub) for sub in wl["subsets"]) * 2 + 4096 * len(wl["linears"])
    st = _h2d_buffers(dev, amax, wmax)
    views, ready, wdev = {}, [], []
    with torch.cuda.stream(cs):
        if st["consumed"][slot] is not None:
            cs.wait_event(st["consumed"][slot])           # the previous occupant of this area is done with it
        off = 0
        for a in ids:
            X = acts[a]
            n = X.numel() * X.element_size()
            views[a] = st["bufs"][slot][off:off + n].view(X.dtype).view(X.shape)
            off += n
        for j in range(n_local):
            for a in ids:
                views[a][j].copy_(acts[a][j], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            ready.append(ev)
        off = 0
        for li in subset:
            w = weights[li]
            nb = w.numel() * w.element_size()
            wd = st["wbufs"][slot][off:off + nb].view(w.dtype).view(w.shape)
            if wrows is not None:                          # row-sharded solve: this rank reads its own rows only (the rest of
                lo, hi = wrows[li]                         # the layer's weight is produced by the all-gather)
                wd[lo:hi].copy_(w[lo:hi], non_blocking=True)
            else:
                wd.copy_(w, non_blocking=True)
            wdev.append(wd)
            off += (nb + 255) // 256 * 256
        wev = torch.cuda.Event()
        wev.record(cs)
    return views, ready, wdev, wev


def quantize_layer(wl, dev, acts, weights, world, from_host, carry=None, prefetch_next=False):
    """One step: the reference driver's per-layer flow (llama.py:97-170 / opt.py:109-214) over our GPTQ / Quantizer / pack."""
    import torch.nn as nn
    import gptq_b200.gptq as G
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer, QuantLinear, pack_codes
    from gptq_b200 import dist as gd
    results = []
    d2h = 0
    cur = torch.cuda.current_stream()
    pending = None
    subsets, linears = wl["subsets"], wl["linears"]
    rank = int(os.environ.get("RANK", "0"))
    wrows = None
    if world > 1:
        wrows = {li: gd.shard_range(linears[li][1], rank, world) for li in range(len(linears))}
    if from_host:
        # `carry`: the previous step already queued this step's first subset behind its own last transfer
        pending = carry if carry is not None else _enqueue_subset_h2d(wl, dev, acts, weights, subsets[0], 0, wrows)
    # the layer's modules exist before the step, like the model's in the reference drivers; only their weights are reset
    mods = _LAYERS.get((id(wl), dev))
    if mods is None:
        mods = []
        for name, r, c, a in linears:
            lin = nn.Linear(c, r, bias=False, device="meta")
            lin.to_empty(device=dev)
            mods.append(lin.half())
        _LAYERS[(id(wl), dev)] = mods
    for si, subset in enumerate(subsets):
        gs = []
        for li in subset:
            name, r, c, a = linears[li]
            lin = mods[li]
            if not from_host:
                lin.weight.data = weights[li].clone()
            g = GPTQ(lin)
            g.quantizer = Quantizer()
            g.quantizer.configure(wl["wbits"], perchannel=True, sym=False, mse=False)
            gs.append(g)
        n_local = acts[0].shape[0]
        if from_host:
            # the H2D stream runs one subset ahead of the compute stream (while this subset's Hessians, chain and
            # column loops run, the next subset's samples are already landing in the other buffer)
            views, ready, wdev, wev = pending
            if si + 1 < len(subsets) or prefetch_next:      # the upload keeps running one subset ahead, across steps too
                nxt = subsets[(si + 1) % len(subsets)]
                pending = _enqueue_subset_h2d(wl, dev, acts, weights, nxt, (si + 1) & 1, wrows)
            else:
                pending = None
            for j in range(n_local):
                cur.wait_event(ready[j])
                for li, g in zip(subset, gs):
                    g.add_batch(views[linears[li][3]][j].unsqueeze(0), None)
            cur.wait_event(wev)                           # this subset's weights have landed
            for g, wd in zip(gs, wdev):
                g.layer.weight.data = wd
        else:
            for j in range(n_local):
                for li, g in zip(subset, gs):             # the forward hooks of q/k/v see the same tensor
                    g.add_batch(acts[linears[li][3]][j].unsqueeze(0), None)
        for g in gs:
            if world > 1:
                gd.allreduce_hessian(g, total_samples=NSAMPLES)
                gd.fasterquant_sharded(g, **wl["gptq"])
            else:
                g.fasterquant(**wl["gptq"])
            with G._prof("pack"):
                if wl["pack"] == "linear":                # the reference's entry point: fp16 weights + [R,ng] scales / zeros
                    ql = QuantLinear(wl["wbits"], g.columns, g.rows)
                    ql.pack(g.layer, g.scales, g.zeros, g_idx=g.g_idx)
                    qw = ql.qweight
                else:
                    qw = pack_codes(g.codes, wl["wbits"])
            if from_host:                                 # result back to pinned host memory: packed words + scales + zeros
                for key, t in (("qw", qw), ("scales", g.scales), ("zeros", g.zeros), ("g_idx", g.g_idx)):
                    d2h += _d2h(dev, (len(results), key), t)  # asynchronous; the step's closing synchronisation covers them
            results.append((qw, g))
            g.free()
        if from_host:                                     # activations staged, weights read: the landing area is free
            ev = torch.cuda.Event()
            ev.record(cur)
            _H2D[dev]["consumed"][si & 1] = ev
    if from_host and pending is None:
        cur.wait_stream(_copy_stream(dev))
    return results, d2h, pending


def multi_gpu_parity(wl, dev, rank, world):
    """Outside the timed region: the FIRST linear of the layer quantized (a) sample- and row-sharded over all ranks and
    (b) on rank 0 alone with all samples; reports how the two results compare (replicas must be bitwise equal,
    the sharded result differs from the single-GPU one only through the summation order of the partial Hessians)."""
    import torch.distributed as dist
    import torch.nn as nn
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer
    from gptq_b200 import dist as gd
    name, R, K, a = wl["linears"][0]
    ns = 32                                                # enough samples for a well-conditioned K x K Hessian
    gw = torch.Generator(device=dev).manual_seed(1)
    W = (torch.randn(R, K, device=dev, generator=gw) * 0.02).half()
    g0 = torch.Generator(device=dev).manual_seed(100 + a)
    chan = torch.exp(0.5 * torch.randn(K, device=dev, generator=g0))

    def sample(j):
        gj = torch.Generator(device=dev).manual_seed(1000 * (a + 1) + j)
        return (torch.randn(1, SEQ, K, device=dev, generator=gj) * chan).half()

    def run(samples, sharded):
        lin = nn.Linear(K, R, bias=False, device="meta")
        lin.to_empty(device=dev)
        lin = lin.half()
        lin.weight.data = W.clone()
        g = GPTQ(lin)
        g.quantizer = Quantizer()
        g.quantizer.configure(wl["wbits"], perchannel=True, sym=False, mse=False)
        for j in samples:
            g.add_batch(sample(j), None)
        with contextlib.redirect_stdout(io.StringIO()):
            if sharded:
                gd.allreduce_hessian(g, total_samples=ns)
                gd.fasterquant_sharded(g, **wl["gptq"])
            else:
                g.fasterquant(**wl["gptq"])
        return g.codes, lin.weight.data, g.error
    codes_s, wq_s, err_s = run(gd.sample_shard(ns, rank, world), True)
    # replicas: a 64-bit checksum of codes and weights, compared across ranks
    chk = torch.stack([codes_s.long().sum(), wq_s.view(torch.int16).long().sum(),
                       (codes_s.long() * torch.arange(1, K + 1, device=dev)).sum()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    replicas_equal = all(torch.equal(c, allc[0]) for c in allc)
    out = None
    if rank == 0:
        codes_1, wq_1, err_1 = run(range(ns), False)
        diff = codes_1 != codes_s
        out = {"linear": f"{name} {R}x{K}", "samples": ns, "replicas_bitwise_equal": bool(replicas_equal),
               "codes_equal": bool(not diff.any()), "codes_agree": float(1.0 - diff.double().mean()),
               "rows_differing": int(diff.any(dim=1).sum()), "weights_agree": float((wq_1 == wq_s).double().mean()),
               "error_single": err_1, "error_sharded": err_s,
               "note": "sharded = partial Hessians summed by the all-reduce (different fp32 summation order than one GPU), rows solved "
                       "on different ranks and gathered; gate: >= 99.9 % codes agree"}
    dist.barrier()
    return out


def run_ours(args):
    import torch.distributed as dist
    from gptq_b200 import _lib
    import gptq_b200.gptq as G

    wl = WORKLOADS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("GPTQ_B200_EXPANDABLE_SEGMENTS", "1") == "1":
        # same allocator setting as the whole-model driver (bench_model.py): the K x K and R x K fp32 buffers
        # of the solver (64 - 485 MB each) otherwise make the caching allocator call cudaMalloc / cudaFree inside steps
        # whenever its block-splitting pattern shifts (seen as single steps with a 50 - 170 ms "prologue")
        torch.cuda.memory._set_allocator_settings("expandable_segments:True")
    numa = numa_bind(local) if world > 1 else {"applied": False, "note": "single GPU: not applied (the CPU baseline of this run uses every host thread)"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    STEP_MS = []
    ALLOC = {}
    STEP_ALLOCS = []

    def timed(acts, weights, from_host, steps, warmup, profile):
        quiet = io.StringIO()
        d2h = 0
        with contextlib.redirect_stdout(quiet):
            G.PROFILE = [] if profile else None        # the warm-up runs exactly what the timed steps run (event pools included)
            # (the warm-up holds the previous step's results while the next step runs, exactly like the timed loop below: when
            # it did not, the first timed steps had a higher memory peak than any warm-up step and the allocator grew its pool
            # inside the timed region -- seven cuMemCreate / cuMemMap calls, each a driver round trip that occasionally
            # collided with the clock sampler's NVML queries and showed up as one step of 200 - 900 ms, always the second)
            _ = None
            for _w in range(warmup):
                _ = quantize_layer(wl, dev, acts, weights, world, from_host)
            carry = None
            # slack in the allocator's pool: whatever the timed steps need beyond the warm-up's peak (block-splitting patterns
            # shift from step to step) is served from memory that is already mapped
            slack = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
            small = [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _s in range(256)]   # the pool of blocks < 1 MB
            del slack, small
            # Python's cyclic garbage collector is kept out of the timed region: with torch / transformers imported a
            # generation-2 pass walks millions of objects and showed up as single steps of 190 - 400 ms (stated in `config`)
            gc.collect()
            gc.freeze()
            gc.disable()
            barrier()
            G.PROFILE = [] if profile else None
            ms0 = torch.cuda.memory_stats(dev)
            n0 = _lib.launch_count()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            marks = []
            for i in range(steps):
                _, d2h, carry = quantize_layer(wl, dev, acts, weights, world, from_host, carry, prefetch_next=from_host and i + 1 < steps)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
                STEP_ALLOCS.append(torch.cuda.memory_stats(dev).get("num_device_alloc", 0))
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            gc.enable()
            gc.unfreeze()
            ms1 = torch.cuda.memory_stats(dev)
            ALLOC.clear()
            ALLOC.update({k: ms1.get(k, 0) - ms0.get(k, 0) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")})
            ALLOC["steps_with_device_alloc"] = [i for i, v in enumerate(STEP_ALLOCS)
                                                if v != (STEP_ALLOCS[i - 1] if i else ms0.get("num_device_alloc", 0))]
            STEP_ALLOCS.clear()
        ms = e0.elapsed_time(e1)
        STEP_MS.clear()
        STEP_MS.extend(round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(steps))
        prof = G.PROFILE
        G.PROFILE = None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, _lib.launch_count() - n0, prof, d2h

    # ---- device-resident run (value)
    # The clock sampler (nvidia-smi in loop mode) is started BEFORE the inputs are generated: its start-up (NVML
    # initialisation, device enumeration) takes driver locks for a few hundred milliseconds, which showed up as one 300 ms
    # step inside the timed region when it was started right in front of it.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    acts, weights = make_inputs(wl, dev, rank, world, host=False)
    if rank == 0:
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.05)
    ms, wall, launches, prof, _ = timed(acts, weights, False, args.steps, args.warmup, True)
    step_ms_each = list(STEP_MS)
    alloc_timed = dict(ALLOC, expandable_segments=os.environ.get("GPTQ_B200_EXPANDABLE_SEGMENTS", "1") == "1")
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    value = wl["weights"] / (ms_step / 1e3) / 1e6

    # per-kernel device time inside the timed region (CUDA events on the launching stream)
    agg = {}
    for name, a, b, work, _meta in prof:
        d = agg.setdefault(name, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += work
        d[2] += 1
    sy = agg.get("syrk", [1e-9, 0.0, 1])
    syrk_tflops = sy[1] / (sy[0] / 1e3) / 1e12
    dense_equiv = sum(2.0 * c * c * (NSAMPLES // world) * SEQ for _, r, c, _ in wl["linears"]) * args.steps / (sy[0] / 1e3) / 1e12
    traffic, traffic_src = syrk_traffic(prof) if world == 1 else (None, None)
    kname = {1: "syrk_tc_kernel (tcgen05 cta_group::1, M128 x N256 per SM)"}.get(G.SYRK_VARIANT, "syrk_tc2_kernel (tcgen05 cta_group::2, M256 x N256 per CTA pair)")
    roofline = {"kernel": kname + " -- Hessian accumulation", "bound": "tensor", "achieved": syrk_tflops,
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": syrk_tflops / pk["tf_sustained"],
                "peak_source": pk["source"] + " bf16 cuBLAS sustained", "frac_of_burst_peak": syrk_tflops / pk["tf_burst"],
                "frac_of_nominal_2250": syrk_tflops / 2250.0,
                "launches": sy[2], "avg_launch_ms": sy[0] / max(sy[2], 1), "share_of_step": sy[0] / ms,
                "dense_equivalent_tflops": dense_equiv, "traffic": traffic,
                "traffic_note": f"DRAM read+write bytes per launch, launch-weighted over the launch shapes of the timed region, from the "
                                f"committed ncu capture profiles/{traffic_src}"}
    co = agg.get("colloop", [1e-9, 0.0, 1])
    hv = agg.get("hinv_upper", [1e-9, 0.0, 1])
    others = {
        "colloop": {"bound": "hbm", "achieved": co[1] / (co[0] / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": co[1] / (co[0] / 1e3) / 1e9 / pk["hbm_gbs"], "share_of_step": co[0] / ms,
                    "us_per_column": co[0] * 1e3 / (args.steps * sum(c for _, r, c, _ in wl["linears"])),
                    "note": "sequential column recurrence: latency-bound, low HBM fraction expected (SURVEY 8d)"},
        "hinv_upper": {"bound": "fp32", "achieved": hv[1] / (hv[0] / 1e3) / 1e12, "unit": "TFLOP/s (2/3 K^3 useful)",
                       "share_of_step": hv[0] / ms},
    }
    phase_ms = {k: v[0] / args.steps for k, v in agg.items()}
    phase_ms["rest (see phase_breakdown_instrumented_step)"] = ms_step - sum(phase_ms.values())

    # ---- one extra, fully instrumented step (NOT part of the timed region): events around every staging copy / verify /
    #      collective / pack, and the host time spent blocked in fasterquant's record read-back
    G.PROFILE, G.PROFILE_FINE, G.HOST_WAIT = [], True, [0.0]
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with contextlib.redirect_stdout(io.StringIO()):
        ea.record()
        quantize_layer(wl, dev, acts, weights, world, False)
        eb.record()
    torch.cuda.synchronize()
    fine = {}
    for name, a, b, work, _meta in G.PROFILE:
        fine[name] = fine.get(name, 0.0) + a.elapsed_time(b)
    fine_total = ea.elapsed_time(eb)
    fine["unattributed"] = fine_total - sum(fine.values())
    fine["_step_ms_with_instrumentation"] = fine_total
    fine["_host_blocked_in_readback_ms"] = G.HOST_WAIT[0] * 1e3
    G.PROFILE, G.PROFILE_FINE, G.HOST_WAIT = None, False, None

    parity = None
    if world > 1:
        parity = multi_gpu_parity(wl, dev, rank, world)
    del acts, weights
    torch.cuda.empty_cache()

    # ---- end-to-end run: host buffers, H2D/D2H inside the timed region
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        acts_h, weights_h = make_inputs(wl, dev, rank, world, host=True)
        h2d = sum(x.numel() * 2 for x in acts_h) + sum(w.numel() * 2 for w in weights_h) // world   # own samples + own weight rows
        e_steps = max(1, min(args.steps, 3))
        ems, _, _, _, d2h = timed(acts_h, weights_h, True, e_steps, 1, False)
        e2e = {"value": wl["weights"] / (ems / e_steps / 1e3) / 1e6, "unit": "Mweights/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": ems / e_steps, "steps": e_steps}
        del acts_h, weights_h
    except Exception as ex:  # pinned allocation can fail on small hosts; report instead of hiding
        if not args.no_e2e:
            import traceback
            traceback.print_exc(file=sys.stderr)
        e2e = {"value": None, "unit": "Mweights/s", "error": repr(ex)}
    _H2D.clear()
    torch.cuda.empty_cache()

    if rank == 0:
        out = base_line(args, wl, value, ms_step)
        out.update({"hot_path_seconds_x_layers": ms_step * wl["n_layers"] / 1e3, "wall_seconds_timed_region": wall,
                    "hessian_tflops_issued": syrk_tflops, "hessian_tflops_dense_equivalent": dense_equiv,
                    "roofline": roofline, "rooflines_other": others, "phase_ms_per_step": phase_ms,
                    "phase_breakdown_instrumented_step": fine, "clocks": clocks, "e2e": e2e,
                    "gpu_launches": launches, "impl": "ours", "ms_each_step_rank0": step_ms_each,
                    "allocator_timed_region_rank0": alloc_timed, "numa_rank0": numa})
        if parity is not None:
            out["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample(wl)
            out["cpu_baseline"] = {"value": wl["weights"] / cb["layer_seconds"] / 1e6, "unit": "Mweights/s", "cores": cb["cores"],
                                   "kind": cb["kind"], "sample": cb["sample"], "detail": cb["measured"],
                                   "extrapolated_layer_seconds": cb["layer_seconds"]}
        if world == 1 and not args.no_gpu_baseline:
            try:
                out["gpu_torch_baseline"] = reference_cuda_layer(wl, dev, 1, 1)
            except Exception as ex:
                out["gpu_torch_baseline"] = {"value": None, "error": repr(ex)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--config", default="llama7b", choices=sorted(WORKLOADS))
    ap.add_argument("--workload", default="layer", choices=["layer", "model"])
    ap.add_argument("--model-layers", type=int, default=0, help="--workload model: quantize only the first N decoder layers (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "reference-cuda":
        run_reference_cuda(args)
        return
    if int(os.environ.get("WORLD_SIZE", "1")) != args.gpus and args.gpus > 1:
        # convenience: spawn torchrun ourselves when launched as plain `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "model":
        from bench_model import run_model
        run_model(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
