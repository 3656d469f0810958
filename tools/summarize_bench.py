"""Copy the bench lines of this round from gpurun_out/ into profiles/ and write profiles/r02_bench_summary.md.

    python tools/summarize_bench.py
"""
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def last_json(path):
    if not os.path.exists(path):
        return None
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except Exception:
                continue
    return None


def keep(src, dst):
    d = last_json(os.path.join(OUT, src))
    if d is not None:
        with open(os.path.join(PROF, dst), "w") as f:
            json.dump(d, f, indent=1)
    return d


lines = ["# Round 2: bench lines (`bench.py` on the GPU box; full JSON beside this file)", ""]
layer = {}
# N = 1, 2, 4, 8 back to back on ONE 8-GPU box (tools/gpu_call_8.sh); the N = 1 line with the CPU / GPU-torch baselines comes from a
# separate 1-GPU box (tools/gpu_call13.sh)
for n, src in ((1, "r2_scale_n1.json"), (2, "r2_scale_n2.json"), (4, "r2_scale_n4.json"), (8, "r2_scale_n8.json")):
    d = keep(src, f"r02_bench_n{n}.json" if n > 1 else "r02_bench_n1_scaling_box.json")
    if d:
        layer[n] = d
full1 = keep("r2_bench_final_n1.json", "r02_bench_n1.json")
if layer:
    lines += ["## Llama-2-7B decoder layer, 4-bit g128 act-order static groups, 128 x 2048 tokens (BASELINE configs[2])", "",
              "| GPUs | ms/step | Mweights/s | e2e ms/step (host buffers) | e2e Mweights/s | Hessian | chain | column loop | prologue | staging+rest | Hessian TFLOP/s issued | frac of 2250 |",
              "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for n, d in sorted(layer.items()):
        ph = d["phase_ms_per_step"]
        rest = d["ms_per_step"] - sum(v for k, v in ph.items() if k in ("syrk", "hinv_upper", "colloop", "prologue"))
        e = d.get("e2e") or {}
        lines.append(f"| {n} | {d['ms_per_step']:.1f} | {d['value']:.0f} | {e.get('ms_per_step', float('nan')):.1f} | {(e.get('value') or float('nan')):.0f} | "
                     f"{ph.get('syrk', 0):.1f} | {ph.get('hinv_upper', 0):.1f} | {ph.get('colloop', 0):.1f} | {ph.get('prologue', 0):.1f} | {rest:.1f} | "
                     f"{d['roofline']['achieved']:.0f} | {d['roofline']['frac_of_nominal_2250']:.3f} |")
    lines.append("")
    if layer.get(1):
        b = layer[1]["value"]
        lines.append("Scaling efficiency (value(N) / (N x value(1)), same box): " + ", ".join(
            f"N={n}: {d['value'] / b / n:.2f}" for n, d in sorted(layer.items())) + ".")
        lines.append("")
    sb = os.path.join(PROF, "r02_bench_n8_second_box.json")
    if os.path.exists(sb):
        d8 = json.load(open(sb))
        lines += [f"A second N = 8 run of the final build on another 8-GPU box (`profiles/r02_bench_n8_second_box.json`, `--steps {d8['steps']} "
                  f"--warmup {d8['warmup']}`):", f"{d8['ms_per_step']:.1f} ms/step ({d8['value']:.0f} Mweights/s; steps "
                  f"{min(d8['ms_each_step_rank0']):.1f}-{max(d8['ms_each_step_rank0']):.1f} ms), e2e {d8['e2e']['ms_per_step']:.1f} ms/step "
                  f"({d8['e2e']['value']:.0f} Mweights/s).  Both 8-GPU boxes are single-NUMA-node VMs (`numa_node = -1`), so the NUMA-local",
                  "placement of the pinned staging buffers (`numa_bind` in bench.py) is a no-op there.", ""]
    d1 = full1
    if d1:
        lines.append(f"N = 1 on a separate 1-GPU box with both baselines in the same run (`profiles/r02_bench_n1.json`): {d1['ms_per_step']:.1f} ms/step, "
                     f"{d1['value']:.0f} Mweights/s, e2e {d1['e2e']['ms_per_step']:.1f} ms/step = {d1['e2e']['value']:.0f} Mweights/s, Hessian "
                     f"{d1['roofline']['achieved']:.0f} TFLOP/s issued.")
        cb, gb = d1.get("cpu_baseline"), d1.get("gpu_torch_baseline")
        if cb:
            lines.append(f"CPU baseline ({cb['kind']}, {cb['cores']} threads, bounded sample): {cb['value']:.2f} Mweights/s "
                         f"({cb.get('extrapolated_layer_seconds', 0):.1f} s per layer).")
        if gb and gb.get("value"):
            lines.append(f"GPU-torch baseline (the unmodified reference classes with CUDA tensors on the same B200, whole layer): "
                         f"{gb['value']:.1f} Mweights/s = {gb['ms_per_step'] / 1e3:.2f} s per layer (Hessian {gb['hessian_ms'] / 1e3:.2f} s, "
                         f"solve {gb['solve_ms'] / 1e3:.2f} s) -> this implementation is {d1['value'] / gb['value']:.0f}x faster device-resident.")
        lines.append(f"Steps of the timed region (rank 0, ms): {d1.get('ms_each_step_rank0')}")
        lines.append(f"Instrumented extra step (events around every staging copy / verification / collective): {json.dumps(d1.get('phase_breakdown_instrumented_step'))}")
        lines.append("")
    for n, d in sorted(layer.items()):
        if d.get("parity"):
            lines.append(f"N = {n} parity (outside the timed region): {json.dumps(d['parity'])}")
    lines.append("")

ref = keep("r2_bench_ref.json", "r02_bench_reference_arm.json")
if ref:
    lines += ["## Reference arm (`bench.py --impl reference`: the unmodified reference, host CPU)", "",
              f"{ref['value']:.2f} Mweights/s, {ref['cpu_baseline']['cores']} threads, kind `{ref['cpu_baseline']['kind']}`; sample step "
              f"{ref['ms_per_step'] / 1e3:.1f} s; layer assembled from unit costs: {ref['extrapolated_layer_seconds']:.1f} s "
              f"(Hessian {ref['cpu_baseline']['hessian_seconds']:.1f} s + solve {ref['cpu_baseline']['solve_seconds']:.1f} s).", ""]

models = {}
for n, src in ((1, "r2_model_n1.json"), (2, "r2_model_n2.json"), (4, "r2_model_n4.json"), (8, "r2_model_n8.json")):
    d = keep(src, f"r02_model_n{n}.json")
    if d:
        models[n] = d
if models:
    lines += ["## Whole-model quantization, random-init Llama-2-7B (32 layers), `bench.py --workload model` (SURVEY 8d definition)", "",
              "| GPUs | whole-model seconds | capture | hooked forwards + add_batch | fasterquant (+ collectives) | re-forward | Hessian kernels | chain | column loop |",
              "|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for n, d in sorted(models.items()):
        p, k = d["phase_seconds_rank0"], d["kernel_seconds_rank0"]
        lines.append(f"| {n} | {d['value']:.2f} | {p.get('capture_s', 0):.2f} | {p.get('forward_hessian_s', 0):.2f} | {p.get('solve_s', 0):.2f} | "
                     f"{p.get('reforward_s', 0):.2f} | {k.get('syrk', 0):.2f} | {k.get('hinv_upper', 0):.2f} | {k.get('colloop', 0):.2f} |")
    lines += ["", "north_star target: Llama-2-7B fully quantized on 8 x B200 in under 60 s.", ""]

others = {}
for cfg in ("opt125m", "opt1.3b", "bloom7b1", "opt66b"):
    d = keep(f"r2_bench_{cfg}.json", f"r02_bench_{cfg}.json")
    if d:
        others[cfg] = d
if others:
    lines += ["## The other BASELINE configs, one decoder layer each, N = 1 (`bench.py --config ...`)", "",
              "| config | ms/step | Mweights/s | layers | hot path x layers (s) | Hessian | chain | column loop | Hessian TFLOP/s issued | e2e Mweights/s | CPU baseline Mweights/s |",
              "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for cfg, d in others.items():
        ph = d["phase_ms_per_step"]
        e = d.get("e2e") or {}
        cb = d.get("cpu_baseline") or {}
        lines.append(f"| {d['config']['workload'][:60]} | {d['ms_per_step']:.1f} | {d['value']:.0f} | - | {d['hot_path_seconds_x_layers']:.2f} | "
                     f"{ph.get('syrk', 0):.1f} | {ph.get('hinv_upper', 0):.1f} | {ph.get('colloop', 0):.1f} | {d['roofline']['achieved']:.0f} | "
                     f"{(e.get('value') or float('nan')):.0f} | {(cb.get('value') or float('nan')):.2f} |")
    lines.append("")

open(os.path.join(PROF, "r02_bench_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
