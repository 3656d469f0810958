"""Per-linear diagnosis of code agreement on a real (random-init) decoder layer: is disagreement between the CUDA
path and the fp32 CPU reference larger than the disagreement between the fp32 and an fp64 run of the reference?"""
import copy, io, contextlib, sys
import torch
sys.path.insert(0, ".")
from gptq_b200.gptq import GPTQ
from gptq_b200.quant import Quantizer
from gptq_b200 import sequential as S
from oracle import gptq_oracle as O
sys.path.insert(0, "tests")
from test_gpu_e2e_model import _build

dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "opt"
m = _build(kind)
seqlen, nsamples = 256, 16
calib = torch.randint(0, 2048, (nsamples, seqlen), generator=torch.Generator().manual_seed(0))
layers, _, _ = S.decoder_parts(m)
inps, kwargs = S.capture_layer_inputs(m, calib, dev)
layer = layers[0].to(dev)
full = S.find_layers(layer)
acts = {n: [] for n in full}
hs = [full[n].register_forward_hook((lambda n: lambda _, i, o: acts[n].append(i[0].data.clone()))(n)) for n in full]
for j in range(nsamples):
    S._call_layer(layer, inps[j].unsqueeze(0), kwargs)
for h in hs: h.remove()

def fq64(W, H, qp, **kw):
    """reference algorithm with the inverse-Cholesky chain in fp64 (everything else fp32)"""
    orig = O.hinv_upper
    def hu(Hm, percdamp):
        return orig(Hm.double(), percdamp).float()
    O.hinv_upper = hu
    try:
        return O.fasterquant(W, H, qp, **kw)
    finally:
        O.hinv_upper = orig

for name, lin in full.items():
    W = lin.weight.data.clone()
    st = O.HessianState(W.shape[1])
    for x in acts[name]:
        O.add_batch(st, x.cpu())
    Hd = st.H.double().clone()
    dead = torch.diag(Hd) == 0
    Hd[dead, dead] = 1
    Hd += torch.eye(Hd.shape[0], dtype=torch.float64) * 0.01 * torch.mean(torch.diag(Hd))
    ev = torch.linalg.eigvalsh(Hd)
    cond = float(ev[-1] / ev[0])
    kw = dict(groupsize=-1)
    qp = O.QParams(bits=4)
    r32 = O.fasterquant(W.cpu(), st.H, qp, **kw)
    r64 = fq64(W.cpu(), st.H, qp, **kw)
    l2 = copy.deepcopy(lin)
    g = GPTQ(l2); g.quantizer = Quantizer(); g.quantizer.configure(4, perchannel=True, sym=False, mse=False)
    for x in acts[name]:
        g.add_batch(x, None)
    Hg = g.H.cpu()
    with contextlib.redirect_stdout(io.StringIO()):
        g.fasterquant(groupsize=-1)
    cg = g.codes.cpu()
    # same-H run of the GPU solver
    l3 = copy.deepcopy(lin)
    g3 = GPTQ(l3); g3.quantizer = Quantizer(); g3.quantizer.configure(4, perchannel=True, sym=False, mse=False)
    g3.add_batch(acts[name][0], None)
    g3.H = st.H.to(dev).clone()
    with contextlib.redirect_stdout(io.StringIO()):
        g3.fasterquant(groupsize=-1)
    print(f"{name:22s} K={W.shape[1]:5d} cond={cond:9.2e} dead={int(dead.sum())} Hrel={O.rel_fro(Hg, st.H):.1e} "
          f"gpu~ref32={1-O.code_agreement(cg, r32.codes):.2e} gpuSameH~ref32={1-O.code_agreement(g3.codes.cpu(), r32.codes):.2e} "
          f"ref64~ref32={1-O.code_agreement(r64.codes, r32.codes):.2e} gpu~ref64={1-O.code_agreement(cg, r64.codes):.2e} "
          f"err gpu {g.error:.4g} ref32 {r32.error:.4g} ref64 {r64.error:.4g}")
