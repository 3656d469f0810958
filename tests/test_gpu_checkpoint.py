"""Checkpoint format I/O (reference opt.py:362-402 opt_pack3 / load_quant3, quant.py:140-150,205-216):
quantize -> pack -> torch.save(state_dict) -> fresh model -> make_quant -> load_state_dict -> single-token forward equals
the dense quantized model.  3-bit per-row (the reference's own format, key for key) and 4-bit g64 act-order static groups
(per-group tables [R, ngroups] + g_idx, which the reference cannot save correctly -- SURVEY 8a row J)."""
import contextlib
import copy
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    import transformers
    torch.manual_seed(0)
    cfg = transformers.OPTConfig(hidden_size=128, ffn_dim=256, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                                 max_position_embeddings=128, word_embed_proj_dim=128)
    return transformers.OPTForCausalLM(cfg).half().eval()


@pytest.mark.parametrize("qa", [dict(wbits=3), dict(wbits=4, groupsize=64, act_order=True, static_groups=True)],
                         ids=["b3_per_row", "b4_g64_act_static"])
def test_save_load_roundtrip_single_token_forward(qa, tmp_path):
    from gptq_b200.gptq import GPTQ
    from gptq_b200 import quant as Q
    from gptq_b200 import sequential as S
    dev = torch.device("cuda:0")
    calib = torch.randint(0, 512, (16, 64), generator=torch.Generator().manual_seed(0))
    dense = _model()
    with contextlib.redirect_stdout(io.StringIO()):
        quantizers, extras = S.quantize_sequential(dense, calib, dev, S.QuantArgs(**qa), GPTQ, Q.Quantizer)
    bits, gs = qa["wbits"], qa.get("groupsize", -1)
    packed = copy.deepcopy(dense)
    Q.pack_model(packed, quantizers, bits, extras=extras, groupsize=gs)
    sd = packed.state_dict()
    some = "model.decoder.layers.0.self_attn.q_proj."
    assert sd[some + "qweight"].dtype == torch.int32 and sd[some + "qweight"].shape == (128 // 32 * bits, 128)
    if gs == -1:
        # exactly the reference's keys and shapes for a packed layer (quant.py:144-149)
        keys = sorted(k[len(some):] for k in sd if k.startswith(some))
        assert keys == ["bias", "qweight", "scales", "zeros"]
        assert sd[some + "scales"].shape == (128, 1) and sd[some + "zeros"].shape == (128, 1) and sd[some + "bias"].shape == (128,)
    else:
        assert sd[some + "scales"].shape == (128, 2) and sd[some + "g_idx"].shape == (128,)
    path = tmp_path / "ckpt.pt"
    torch.save(sd, path)

    fresh = Q.load_quant(_model(), str(path), bits, groupsize=gs)
    qlayers = Q.find_linears(fresh, (Q.QuantLinear,))
    assert len(qlayers) == 12 and "lm_head" not in qlayers
    for k, v in fresh.state_dict().items():
        assert torch.equal(v.cpu(), sd[k].cpu()), k

    # single-token forward (the only shape the packed matvec supports, quant.py:203) through the WHOLE model
    dense = dense.to(dev)
    fresh = fresh.to(dev)
    for tok in (3, 77, 300):
        ids = torch.tensor([[tok]], device=dev)
        with torch.no_grad():
            want = dense(ids).logits.float()
            got = fresh(ids).logits.float()
        rel = float((got - want).norm() / want.norm())
        print(f"  [metric] checkpoint roundtrip {qa}: token {tok} logits rel diff {rel:.2e}")
        assert rel < 2e-2            # fp16 model; the packed matvec accumulates in fp32, the dense layer in fp16
    # the packed integer codes are the ones the quantizer produced
    name = "model.decoder.layers.1.fc1"
    codes = Q.unpack_codes(qlayers[name].qweight.to(dev), bits)
    assert torch.equal(codes.cpu(), extras[name]["codes"].cpu())


def test_reference_format_checkpoint_loads(tmp_path):
    """A checkpoint written in the reference's per-row 3-bit format (words from the oracle's restatement of
    Quant3Linear.pack, quant.py:152-187) loads into a make_quant3'ed layer and multiplies like the dense layer."""
    from oracle import gptq_oracle as O
    from gptq_b200 import quant as Q
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    R, K = 96, 160
    codes = torch.randint(0, 8, (R, K), generator=g)
    scales = torch.rand(R, 1, generator=g) * 0.05 + 0.01
    zeros = torch.randint(0, 8, (R, 1), generator=g).float()
    W = (scales * (codes.float() - zeros))
    sd = {"qweight": torch.from_numpy(O.pack3(np.ascontiguousarray(codes.numpy().T).astype(np.uint32))),
          "scales": scales, "zeros": zeros * scales, "bias": torch.zeros(R)}
    torch.save(sd, tmp_path / "ref.pt")
    layer = Q.Quant3Linear(K, R)
    layer.load_state_dict(torch.load(tmp_path / "ref.pt"))
    x = torch.randn(K, generator=g)
    y = layer.to(dev)(x.to(dev).reshape(1, 1, K)).reshape(-1).cpu()
    want = W.double() @ x.double()
    assert float((y.double() - want).norm() / want.norm()) < 1e-5
