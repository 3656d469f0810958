"""How much does the tensor-core accumulator truncation (syrk chunk length) move GPTQ's codes?
Runs fasterquant on a Llama-sized linear with H accumulated at several chunk lengths and compares codes
against the chunk=64 run (fp32-level H) and against the CPU oracle."""
import json
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, ".")
import gptq_b200.gptq as G
from gptq_b200.quant import Quantizer
from oracle import gptq_oracle as O

DEV = torch.device("cuda:0")
R, K, NS, SEQ = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 4096, 16, 2048
cfgs = [dict(name="b4_g128_act_static", bits=4, groupsize=128, actorder=True, static_groups=True),
        dict(name="b3_g128_dynamic", bits=3, groupsize=128)]
g0 = torch.Generator().manual_seed(0)
chan = torch.exp(0.5 * torch.randn(K, generator=g0))
X = [(torch.randn(1, SEQ, K, generator=g0) * chan).half() for _ in range(NS)]
W = (torch.randn(R, K, generator=g0) * 0.02).half()


def run(cfg, chunk):
    G.CHUNK_TOKENS = chunk
    lin = nn.Linear(K, R, bias=False).half().to(DEV)
    lin.weight.data = W.to(DEV).clone()
    g = G.GPTQ(lin)
    g.quantizer = Quantizer()
    g.quantizer.configure(cfg["bits"], perchannel=True, sym=False, mse=False)
    for x in X:
        g.add_batch(x.to(DEV), None)
    H = g.H.cpu()
    g.fasterquant(groupsize=cfg["groupsize"], actorder=cfg.get("actorder", False), static_groups=cfg.get("static_groups", False))
    return g.codes.cpu(), H, g.error


t0 = time.time()
st = O.HessianState(K)
for x in X:
    O.add_batch(st, x)
print(json.dumps(dict(cpu_hessian_s=time.time() - t0)))
for cfg in cfgs:
    t0 = time.time()
    ref = O.fasterquant(W, st.H, O.QParams(bits=cfg["bits"]), groupsize=cfg["groupsize"], actorder=cfg.get("actorder", False),
                        static_groups=cfg.get("static_groups", False))
    t_cpu = time.time() - t0
    base = None
    for chunk in (64, 128, 256, 512, 2048):
        codes, H, err = run(cfg, chunk)
        if base is None:
            base = codes
        print(json.dumps(dict(case=cfg["name"], R=R, K=K, chunk=chunk, H_rel_fro_vs_cpu=O.rel_fro(H, st.H),
                              mismatch_vs_cpu=1 - O.code_agreement(codes, ref.codes),
                              mismatch_vs_chunk64=1 - O.code_agreement(codes, base), err=err, err_cpu=ref.error, cpu_s=t_cpu)))
