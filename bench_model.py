"""`bench.py --workload model`: WHOLE-MODEL quantization time (SURVEY 8d definition: what the reference prints as "full
quantization time", opt.py:686-691 = calibration capture + per-layer hooked forwards + all add_batch + all fasterquant +
the re-forward through every quantized layer; model construction and evaluation excluded).

Random-init HF model of the config's published dimensions (there is no network for checkpoints), synthetic calibration
tokens `randint(0, vocab, (128, 2048))`, seed 0.  The model lives on the GPU (the reference keeps it on the
host and moves one decoder layer at a time; the 0.4 GB per layer over PCIe is not what is being measured here).
N > 1: every rank holds the model and forwards its own calibration samples (gptq_b200.sequential, sample-sharded).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import time

import torch


def build_model(name: str, dev, layers: int = 0):
    import transformers
    torch.manual_seed(0)
    if name == "llama7b":
        cfg = transformers.LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                                       num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096)
        cls = transformers.LlamaForCausalLM
    elif name in ("opt125m", "opt1.3b"):
        h, f, n, a = (768, 3072, 12, 12) if name == "opt125m" else (2048, 8192, 24, 32)
        cfg = transformers.OPTConfig(hidden_size=h, ffn_dim=f, num_hidden_layers=n, num_attention_heads=a, vocab_size=50272,
                                     max_position_embeddings=2048, word_embed_proj_dim=h)
        cls = transformers.OPTForCausalLM
    elif name == "bloom7b1":
        cfg = transformers.BloomConfig(hidden_size=4096, n_layer=30, n_head=32, vocab_size=250880)
        cls = transformers.BloomForCausalLM
    else:
        raise ValueError(f"--workload model is not defined for --config {name}")
    if layers:
        for key in ("num_hidden_layers", "n_layer"):
            if hasattr(cfg, key):
                setattr(cfg, key, layers)
    cfg.use_cache = False
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float16)
    try:
        with torch.device(dev):
            model = cls(cfg)
    finally:
        torch.set_default_dtype(old)
    model.eval()
    model.seqlen = 2048
    return model, cfg


MODEL_QUANT = {
    "llama7b": dict(wbits=4, groupsize=128, act_order=True, static_groups=True, true_sequential=True),
    "opt125m": dict(wbits=4, groupsize=-1, static_groups=True),
    "opt1.3b": dict(wbits=3, groupsize=128, static_groups=True),
    "bloom7b1": dict(wbits=2, groupsize=64),
}


def run_model(args):
    import torch.distributed as dist
    from gptq_b200 import _lib
    import gptq_b200.gptq as G
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer
    from gptq_b200 import sequential as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("GPTQ_B200_EXPANDABLE_SEGMENTS", "1") == "1":
        # The solver's K x K and R x K buffers (64 - 485 MB each) alternate with the model forward's activations; with the
        # default allocator the forward splits the cached big blocks and every layer pays ~150 cudaMalloc calls per 6 layers
        # (measured: prologue 56 ms/layer instead of 3.7, 5.0 s instead of 4.0 s for 6 layers).  Expandable segments avoid it;
        # drivers should run with PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True.
        torch.cuda.memory._set_allocator_settings("expandable_segments:True")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    nsamples, seqlen = 128, 2048
    model, cfg = build_model(args.config, dev, args.model_layers)
    vocab = cfg.vocab_size
    tokens = torch.randint(0, vocab, (nsamples, seqlen), generator=torch.Generator().manual_seed(0))
    qa = S.QuantArgs(**MODEL_QUANT[args.config])
    n_layers = len(S.decoder_parts(model)[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on a 2-layer slice of the same model would change its weights; instead warm the kernels on a scratch copy of
    # the first decoder layer's linears (allocator pools, cuBLAS handles of the forward, NCCL communicators)
    with contextlib.redirect_stdout(io.StringIO()):
        warm, _ = build_model(args.config, dev, 1)
        S.quantize_sequential(warm, tokens[: 2 * world], dev, qa, GPTQ, Quantizer, rank=rank, world=world, offload=False,
                              early_exit=True)
        del warm
    torch.cuda.empty_cache()
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()          # keep generation-2 collections (100+ ms with transformers imported) out of the timed region
    barrier()
    stats = {}
    ms0 = torch.cuda.memory_stats(dev)
    G.PROFILE = []
    n0 = _lib.launch_count()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        quantizers, extras = S.quantize_sequential(model, tokens, dev, qa, GPTQ, Quantizer, rank=rank, world=world, offload=False,
                                                   early_exit=True, stats=stats)
    barrier()
    wall = time.perf_counter() - t0
    gc.enable()
    prof, G.PROFILE = G.PROFILE, None
    ms1 = torch.cuda.memory_stats(dev)
    alloc = {k: ms1.get(k, 0) - ms0.get(k, 0) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    alloc["reserved_gb"] = ms1.get("reserved_bytes.all.peak", 0) / 1e9
    t = torch.tensor([wall], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    agg = {}
    for name, a, b, work, _meta in prof:
        agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / 1e3
    err = sum(float(e["error"]) for e in extras.values() if e.get("error") is not None)
    if rank == 0:
        weights = sum(int(e["codes"].numel()) for e in extras.values())
        out = {"metric": f"{args.config}_gptq_whole_model_quantize_seconds", "value": wall, "unit": "s", "n_gpus": world, "steps": 1,
               "warmup": 1, "ms_per_step": wall * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
               "dtype": "f16 forward; f16xf16->f32 (Hessian), f32 (solver)", "data": "synthetic", "impl": "ours",
               "config": {"workload": f"whole-model GPTQ of a random-init {args.config} ({n_layers} decoder layers) through "
                                      f"gptq_b200.sequential.quantize_sequential: {MODEL_QUANT[args.config]}, {nsamples}x{seqlen} "
                                      "calibration tokens; timed: capture + hooked forwards + add_batch + fasterquant + re-forward "
                                      "(opt.py:686-691); model resident on the GPU, hooked passes stop after the subset's last linear",
                          "layers_quantized": n_layers, "weights_quantized": weights},
               "phase_seconds_rank0": stats, "kernel_seconds_rank0": agg, "sum_of_layer_errors": err,
               "allocator_rank0": dict(alloc, expandable_segments=os.environ.get("GPTQ_B200_EXPANDABLE_SEGMENTS", "1") == "1"),
               "mweights_per_s": weights / wall / 1e6, "gpu_launches": _lib.launch_count() - n0,
               "target": "north_star: Llama-2-7B fully quantized on 8 x B200 in under 60 s"}
        if args.model_layers:
            out["extrapolated_full_model_seconds"] = (wall - stats.get("capture_s", 0.0)) * (
                {"llama7b": 32, "opt125m": 12, "opt1.3b": 24, "bloom7b1": 30}[args.config] / n_layers) + stats.get("capture_s", 0.0)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
