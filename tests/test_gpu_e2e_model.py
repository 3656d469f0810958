"""End-to-end parity through the reference's driver flow (BASELINE north_star: perplexity on the synthetic eval
stream within 0.5 %, codes >= 99.9 %): a random-init HF decoder is quantized layer by layer twice on identical
activations -- once with the CUDA path, once with the CPU oracle -- and evaluated like opt_eval (opt.py:316-333)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(kind):
    import transformers
    torch.manual_seed(0)
    if kind == "opt":
        cfg = transformers.OPTConfig(hidden_size=256, ffn_dim=1024, num_hidden_layers=3, num_attention_heads=8, vocab_size=2048,
                                     max_position_embeddings=512, word_embed_proj_dim=256)
        m = transformers.OPTForCausalLM(cfg)
    else:
        cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=3, num_attention_heads=8,
                                       num_key_value_heads=8, vocab_size=2048, max_position_embeddings=512)
        m = transformers.LlamaForCausalLM(cfg)
    return m.half().eval()


@pytest.mark.parametrize("kind,qa", [
    ("opt", dict(wbits=4)),                                                        # BASELINE config 1 (OPT, per-channel), reduced width
    ("opt", dict(wbits=3, groupsize=128)),                                         # config 2
    ("llama", dict(wbits=4, groupsize=128, act_order=True, static_groups=True, true_sequential=True)),   # config 3
])
def test_sequential_quantization_matches_oracle(kind, qa):
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer
    from gptq_b200 import sequential as S
    from oracle.ref_gptq import OracleGPTQ, OracleQuantizer

    dev = torch.device("cuda:0")
    seqlen, nsamples = 256, 16
    g = torch.Generator().manual_seed(0)
    calib = torch.randint(0, 2048, (nsamples, seqlen), generator=g)
    evalt = torch.randint(0, 2048, (1, 8 * seqlen), generator=torch.Generator().manual_seed(1))
    base = _build(kind)
    args = S.QuantArgs(**qa)

    m_gpu = copy.deepcopy(base)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        _, ex_gpu = S.quantize_sequential(m_gpu, calib, dev, args, GPTQ, Quantizer)
    m_ref = copy.deepcopy(base)
    _, ex_ref = S.quantize_sequential(m_ref, calib, dev, args, OracleGPTQ, OracleQuantizer)

    ppl_fp = S.eval_ppl(copy.deepcopy(base), evalt, dev, seqlen)
    ppl_gpu = S.eval_ppl(m_gpu, evalt, dev, seqlen)
    ppl_ref = S.eval_ppl(m_ref, evalt, dev, seqlen)
    # Codes are comparable only where both runs saw IDENTICAL inputs, i.e. in decoder layer 0: from layer 1 on each run
    # calibrates on the outputs of its own quantized layers, so one legitimate rounding-tie flip upstream changes H.
    # Every differing row of layer 0 must start at a rounding tie of the oracle run (O.classify_mismatches); within the
    # layer, a tie flip in an earlier subset perturbs the later subsets' Hessians (true-sequential): only the linears of the
    # FIRST subset (which see bit-identical inputs in both runs) are classified, with the gpu_H tolerance of
    # tests/test_gpu_headline.py; the later ones are covered by the perplexity gate.
    from oracle import gptq_oracle as O
    tot = mism = tot_all = mism_all = rows_bad = unexplained = 0
    for k in ex_ref:
        a, b = ex_gpu[k]["codes"].cpu(), ex_ref[k]["codes"]
        tot_all += a.numel()
        mism_all += int((a != b).sum())
        first_subset = (not qa.get("true_sequential")) or any(k.endswith(n) for n in S.LLAMA_SEQUENTIAL[0])
        if ".layers.0." in k and first_subset:
            cls = O.classify_mismatches(a, ex_ref[k]["oracle_result"], 2.0 ** -14, perm=ex_gpu[k].get("perm"),
                                        diag=ex_ref[k]["oracle_diag"], perm_tol=1e-5)
            tot += a.numel()
            mism += int((a != b).sum())
            rows_bad += cls["rows_bad"]
            unexplained += cls["unexplained_rows"]
    print(f"  [metric] e2e {kind} {qa}: ppl fp16 {ppl_fp:.3f} gpu {ppl_gpu:.3f} ref {ppl_ref:.3f} "
          f"layer-0 code mismatch {mism / tot:.2e} ({rows_bad} rows, {unexplained} not at a rounding tie); all layers {mism_all / tot_all:.2e}")
    assert abs(ppl_gpu - ppl_ref) <= 5e-3 * ppl_ref
    assert unexplained == 0
