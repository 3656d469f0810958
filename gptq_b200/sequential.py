"""Layer-wise calibration driver: the CALLER of the hot path (SURVEY 8f "next" #2 and #4).

Mirrors ``opt_sequential`` / ``llama_sequential`` / ``bloom_sequential`` (reference opt.py:29-228,
llama.py:31-207, bloom.py:25-187) and ``*_eval`` (opt.py:230-359) for HF decoder-only models, written once
over a small adapter instead of three copies:

* capture the inputs of decoder layer 0 with a ``Catcher`` (opt.py:46-66);
* per layer: wrap every nn.Linear in a ``GPTQ``, hook ``add_batch`` (opt.py:177-187), run the samples,
  ``fasterquant`` each linear, re-run the samples through the quantized layer (opt.py:216-217);
* ``true_sequential`` follows UPSTREAM semantics -- every subset is quantized in order; the fork's
  llama.py:106-110 only quantizes the last subset because of an indentation slip.

The GPTQ / Quantizer classes are parameters, so the same driver runs the CUDA path and (in tests) the
CPU oracle on identical activations.  Works with transformers 5.x (decoder layers take
``position_embeddings``): all keyword arguments seen by layer 0 are replayed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn


@dataclass
class QuantArgs:
    wbits: int = 4
    groupsize: int = -1
    sym: bool = False
    percdamp: float = .01
    act_order: bool = False
    static_groups: bool = False
    true_sequential: bool = False
    trits: bool = False
    nearest: bool = False            # RTN baseline (opt.py:292-300) instead of GPTQ


def find_layers(module, layers=(nn.Conv2d, nn.Linear), name=""):
    """modelutils.py:8-16."""
    if type(module) in layers:
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_layers(child, layers=layers, name=name + "." + name1 if name != "" else name1))
    return res


def decoder_parts(model):
    """(layers ModuleList, modules that must sit on the device while capturing, final norm / head pieces)."""
    if hasattr(model, "transformer") and hasattr(model.transformer, "h"):          # BLOOM (bloom.py:25-60)
        t = model.transformer
        return t.h, [t.word_embeddings, t.word_embeddings_layernorm], [t.ln_f]
    m = model.model
    if hasattr(m, "decoder"):                      # OPT
        d = m.decoder
        pre = [d.embed_tokens, d.embed_positions] + [x for x in (getattr(d, "project_in", None),) if x is not None]
        post = [x for x in (getattr(d, "final_layer_norm", None), getattr(d, "project_out", None)) if x is not None]
        return d.layers, pre, post
    pre = [m.embed_tokens] + [x for x in (getattr(m, "rotary_emb", None),) if x is not None]   # Llama / Gemma
    return m.layers, pre, [m.norm]


def layer_prefix(model) -> str:
    if hasattr(model, "transformer") and hasattr(model.transformer, "h"):
        return "transformer.h"
    return "model.decoder.layers" if hasattr(model.model, "decoder") else "model.layers"


class _StopForward(Exception):
    """Raised by the last forward hook of a subset: the rest of the decoder layer is not needed for the Hessians."""


LLAMA_SEQUENTIAL = [["self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj"], ["self_attn.o_proj"],
                    ["mlp.up_proj", "mlp.gate_proj"], ["mlp.down_proj"]]          # llama.py:98-103


def _call_layer(layer, x, kwargs):
    out = layer(x, **kwargs)
    return out[0] if isinstance(out, (tuple, list)) else out


@torch.no_grad()
def capture_layer_inputs(model, tokens: torch.Tensor, dev, offload: bool = True):
    """opt.py:37-76: run the embeddings on `dev`, stop at decoder layer 0, return (inps [N,S,H], kwargs).
    `offload=False`: the model already lives on `dev` and stays there (no CPU round trips)."""
    layers, pre, _ = decoder_parts(model)
    use_cache = model.config.use_cache
    model.config.use_cache = False
    for p in pre:
        p.to(dev)
    layers[0] = layers[0].to(dev)
    cache: Dict = {"inps": [], "kwargs": None}

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, **kwargs):
            cache["inps"].append(inp.detach())
            cache["kwargs"] = kwargs
            raise ValueError

    layers[0] = Catcher(layers[0])
    for j in range(tokens.shape[0]):
        try:
            model(tokens[j:j + 1].to(dev))
        except ValueError:
            pass
    layers[0] = layers[0].module
    if offload:
        layers[0] = layers[0].cpu()
        for p in pre:
            p.cpu()
    model.config.use_cache = use_cache
    kwargs = {k: v for k, v in cache["kwargs"].items() if k not in ("past_key_values", "past_key_value", "use_cache")}
    kwargs["use_cache"] = False
    return torch.cat(cache["inps"], 0), kwargs


@torch.no_grad()
def quantize_sequential(model, tokens: torch.Tensor, dev, args: QuantArgs, gptq_cls, quantizer_cls,
                        quantize_fn: Optional[Callable] = None, log: Callable = lambda *a: None,
                        rank: int = 0, world: int = 1, group=None, offload: bool = True, early_exit: bool = False,
                        stats: Optional[dict] = None, max_layers: int = 0):
    """The `*_sequential` loop.  Returns {layer-qualified linear name: quantizer} like the reference.

    Multi-GPU (`world` > 1, one process per GPU): every rank holds the model, forwards ITS calibration samples
    (round-robin shard of `tokens`), the partial Hessians are summed by one all-reduce per distinct Hessian, every linear is
    solved row-sharded and completed by one all-gather, and every rank re-forwards its own samples through the quantized
    layer -- so the activations never leave the rank that produced them.
    `early_exit`: a hooked pass stops once every linear of the subset has seen its input (the layer's own output is only
    needed from the final pass, opt.py:216-217, which always runs in full) -- same results, fewer forward flops.
    `stats`: if given, filled with wall-clock seconds per phase (a device synchronisation is inserted at the phase
    boundaries when it is not None)."""
    import time
    from . import dist as gd
    layers, _, _ = decoder_parts(model)
    if world > 1:
        tokens = tokens[gd.sample_shard(tokens.shape[0], rank, world)]
    n_total = None

    def tick():
        if stats is not None and torch.cuda.is_available():
            torch.cuda.synchronize()
        return time.perf_counter()

    def add(name, t0):
        if stats is not None:
            stats[name] = stats.get(name, 0.0) + (tick() - t0)

    t0 = tick()
    inps, kwargs = capture_layer_inputs(model, tokens, dev, offload=offload)
    add("capture_s", t0)
    outs = torch.zeros_like(inps)
    nsamples = inps.shape[0]
    if world > 1:
        cnt = torch.tensor([nsamples], device=inps.device, dtype=torch.int64)
        gd.allreduce_sum_(cnt, group)
        n_total = int(cnt.item())
    quantizers = {}
    extras = {}
    prefix = layer_prefix(model)
    n_layers = len(layers) if not max_layers else min(max_layers, len(layers))
    for i in range(n_layers):
        layer = layers[i].to(dev)
        full = find_layers(layer)
        if args.true_sequential and all(n in full for s in LLAMA_SEQUENTIAL for n in s):
            sequential = LLAMA_SEQUENTIAL
        else:
            sequential = [list(full.keys())]
        for names in sequential:
            subset = {n: full[n] for n in names}
            gptq = {}
            for name in subset:
                gptq[name] = gptq_cls(subset[name])
                gptq[name].quantizer = quantizer_cls()
                gptq[name].quantizer.configure(args.wbits, perchannel=True, sym=args.sym, mse=False, trits=args.trits)
            if args.nearest:                                            # opt.py:292-300
                for name in subset:
                    q = gptq[name].quantizer
                    W = subset[name].weight.data
                    q.find_params(W.float(), weight=True)
                    subset[name].weight.data = quantize_fn(W.float(), q.scale, q.zero, q.maxq).to(W.dtype)
                    quantizers[f"{prefix}.{i}.{name}"] = q
                continue

            fired = [0]

            def add_batch(name):
                def tmp(_, inp, out):
                    gptq[name].add_batch(inp[0].data, out.data)
                    fired[0] += 1
                    if early_exit and fired[0] == len(subset):
                        raise _StopForward
                return tmp
            handles = [subset[name].register_forward_hook(add_batch(name)) for name in subset]
            t0 = tick()
            for j in range(nsamples):
                fired[0] = 0
                try:
                    outs[j] = _call_layer(layer, inps[j].unsqueeze(0), kwargs)[0]
                except _StopForward:
                    pass
            add("forward_hessian_s", t0)
            for h in handles:
                h.remove()
            t0 = tick()
            for name in subset:
                log(i, name)
                kw = dict(blocksize=128, percdamp=args.percdamp, groupsize=args.groupsize, actorder=args.act_order,
                          static_groups=args.static_groups, layer_name=f"{i}.{name}")
                if world > 1:
                    gd.allreduce_hessian(gptq[name], group, total_samples=n_total)
                    gd.fasterquant_sharded(gptq[name], group, **kw)
                else:
                    gptq[name].fasterquant(**kw)
                quantizers[f"{prefix}.{i}.{name}"] = gptq[name].quantizer
                extras[f"{prefix}.{i}.{name}"] = {k: getattr(gptq[name], k, None) for k in ("codes", "scales", "zeros", "g_idx", "error", "perm", "oracle_result", "oracle_diag")}
                gptq[name].free()
            add("solve_s", t0)
        t0 = tick()
        for j in range(nsamples):
            outs[j] = _call_layer(layer, inps[j].unsqueeze(0), kwargs)[0]
        add("reforward_s", t0)
        if offload:
            layers[i] = layer.cpu()
        del layer
        inps, outs = outs, inps
    return quantizers, extras


@torch.no_grad()
def eval_ppl(model, tokens: torch.Tensor, dev, seqlen: int) -> float:
    """opt.py:230-336: layer-wise perplexity of `tokens` ([1, n*seqlen]) -- exp(sum nll / (n*seqlen))."""
    layers, _, post = decoder_parts(model)
    n = tokens.numel() // seqlen
    toks = tokens[:, : n * seqlen].reshape(n, seqlen)
    inps, kwargs = capture_layer_inputs(model, toks, dev)
    for i in range(len(layers)):
        layer = layers[i].to(dev)
        for j in range(n):
            inps[j] = _call_layer(layer, inps[j].unsqueeze(0), kwargs)[0]
        layers[i] = layer.cpu()
    for p in post:
        p.to(dev)
    model.lm_head.to(dev)
    nlls = []
    for j in range(n):
        h = inps[j].unsqueeze(0)
        for p in post:
            h = p(h)
        logits = model.lm_head(h)
        shift_logits = logits[:, :-1, :].float()
        shift_labels = toks[j:j + 1, 1:].to(dev)
        loss = nn.functional.cross_entropy(shift_logits.reshape(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
        nlls.append(loss.float() * seqlen)                              # opt.py:326-331
    for p in post:
        p.cpu()
    model.lm_head.cpu()
    return float(torch.exp(torch.stack(nlls).sum() / (n * seqlen)).item())
