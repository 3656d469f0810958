"""The sample-sharded calibration driver (gptq_b200.sequential.quantize_sequential) on CPU with gloo, world size 2, against
the same driver on one rank: sample sharding, the Hessian all-reduce, the row-sharded solve + all-gather and the per-rank
re-forward are exercised with a small deterministic stand-in for the GPTQ class (the CUDA path itself needs a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class _Q:
    def __init__(self):
        self.scale, self.zero = torch.zeros(1), torch.zeros(1)

    def configure(self, bits, **kw):
        self.maxq = torch.tensor(2 ** bits - 1)


class _FakeGPTQ:
    """Same host-side contract as gptq_b200.gptq.GPTQ (add_batch / raw_accumulator / fasterquant(row_range) / Wq, codes,
    scales, zeros, error_tensor); the 'solver' is a Hessian-diagonal-aware round-to-nearest, rows independent."""

    def __init__(self, layer):
        self.layer = layer
        self.rows, self.columns = layer.weight.shape
        self.S = torch.zeros((self.columns, self.columns), dtype=torch.float64)
        self.nsamples = 0
        self.quantizer = None

    def add_batch(self, inp, out):
        x = inp.reshape(-1, self.columns).double()
        self.S += x.t() @ x
        self.nsamples += 1

    def raw_accumulator(self):
        return self.S

    def fasterquant(self, row_range=None, **kw):
        lo, hi = row_range if row_range is not None else (0, self.rows)
        H = (torch.tril(self.S) + torch.tril(self.S, -1).t()) * (2.0 / self.nsamples)      # consumers mirror the lower triangle
        d = torch.diag(H).clamp_min(1e-9)
        W = self.layer.weight.data[lo:hi].double()
        scale = (W.abs().amax(dim=1, keepdim=True) / 7).clamp_min(1e-12)
        codes = torch.clamp(torch.round(W / scale) + 8, 0, 15)
        Wq = scale * (codes - 8)
        self.Wq, self.codes = Wq.to(self.layer.weight.dtype), codes.to(torch.uint8)
        self.scales, self.zeros = scale.float(), torch.full_like(scale, 8).float()
        self._e = (((W - Wq) ** 2) * d).sum().reshape(1)
        self.error = float(self._e)
        self.quantizer.scale, self.quantizer.zero = self.scales.clone(), self.zeros.clone()
        if row_range is None:
            self.layer.weight.data = self.Wq

    def error_tensor(self):
        return self._e

    def free(self):
        self.S = None


def _model():
    import transformers
    torch.manual_seed(0)
    cfg = transformers.OPTConfig(hidden_size=32, ffn_dim=64, num_hidden_layers=2, num_attention_heads=2, vocab_size=128,
                                 max_position_embeddings=64, word_embed_proj_dim=32)
    return transformers.OPTForCausalLM(cfg).eval()


def _run(rank, world):
    from gptq_b200 import sequential as S
    model = _model()
    tokens = torch.randint(0, 128, (6, 16), generator=torch.Generator().manual_seed(1))
    stats = {}
    S.quantize_sequential(model, tokens, torch.device("cpu"), S.QuantArgs(wbits=4), _FakeGPTQ, _Q, rank=rank, world=world,
                          offload=False, early_exit=(world > 1), stats=stats)
    assert {"capture_s", "forward_hessian_s", "solve_s", "reforward_s"} <= set(stats)
    return {k: v.clone() for k, v in model.state_dict().items() if "layers" in k and k.endswith("weight")}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.save(_run(rank, world), f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_driver_matches_single_rank(tmp_path):
    single = _run(0, 1)
    out = str(tmp_path / "sd")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    changed = 0
    for k in single:
        assert torch.equal(r0[k], r1[k]), k                    # replicas identical after every layer
        assert torch.allclose(r0[k], single[k], atol=1e-6), k  # sharded == one rank (up to the all-reduce's summation order)
        changed += int(not torch.equal(single[k], _model().state_dict()[k]))
    assert changed >= 12                                       # the linears really were quantized


def test_bloom_adapter_finds_layers():
    import transformers
    from gptq_b200 import sequential as S
    cfg = transformers.BloomConfig(hidden_size=32, n_layer=2, n_head=2, vocab_size=64)
    m = transformers.BloomForCausalLM(cfg).eval()
    layers, pre, post = S.decoder_parts(m)
    assert len(layers) == 2 and S.layer_prefix(m) == "transformer.h"
    names = set(S.find_layers(layers[0]))
    assert {"self_attention.query_key_value", "self_attention.dense", "mlp.dense_h_to_4h", "mlp.dense_4h_to_h"} <= names
    tokens = torch.randint(0, 64, (3, 8), generator=torch.Generator().manual_seed(0))
    q, ex = S.quantize_sequential(m, tokens, torch.device("cpu"), S.QuantArgs(wbits=4), _FakeGPTQ, _Q, offload=False)
    assert len(q) == 8 and all(k.startswith("transformer.h.") for k in q)
