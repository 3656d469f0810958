"""Debug: the bench's host-prefetch loop (two alternating device buffers) with shared accumulators."""
import contextlib, io, sys
import torch, torch.nn as nn
sys.path.insert(0, ".")
import gptq_b200.gptq as G
from gptq_b200.gptq import GPTQ
from gptq_b200.quant import Quantizer

dev = torch.device("cuda:0")
K, SEQ, N, R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 2048, 40, 256
g0 = torch.Generator(device=dev).manual_seed(0)
chan = torch.exp(0.5 * torch.randn(K, device=dev, generator=g0))
X = torch.empty((N, SEQ, K), dtype=torch.float16, pin_memory=True)
for j in range(N):
    X[j].copy_((torch.randn(SEQ, K, device=dev, generator=g0) * chan).half())
Href = torch.zeros(K, K, device=dev)
for j in range(N):
    x = X[j].to(dev).float()
    Href += x.t() @ x
Href *= 2.0 / N

def run(share, members=3):
    G.SHARE_H = share
    gs = []
    for i in range(members):
        lin = nn.Linear(K, R, bias=False, device=dev).half()
        g = GPTQ(lin); g.quantizer = Quantizer(); g.quantizer.configure(4, perchannel=True, sym=False, mse=False)
        gs.append(g)
    cs = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream()
    bufs = [torch.empty(X.shape[1:], dtype=X.dtype, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    with torch.cuda.stream(cs):
        bufs[0].copy_(X[0], non_blocking=True)
        ready[0].record(cs)
    for j in range(N):
        b = j & 1
        if j + 1 < N:
            nb = (j + 1) & 1
            with torch.cuda.stream(cs):
                if j >= 1:
                    cs.wait_event(freed[nb])
                bufs[nb].copy_(X[j + 1], non_blocking=True)
                ready[nb].record(cs)
        cur.wait_event(ready[b])
        xj = bufs[b].unsqueeze(0)
        for g in gs:
            g.add_batch(xj, None)
        freed[b].record(cur)
    torch.cuda.synchronize()
    for i, g in enumerate(gs):
        acc = g._acc
        print(f"share={share} member {i}: nsamples={g.nsamples} acc.count={acc.count} members={acc.members} batches={acc.batches} "
              f"staged={acc.staged} flag={None if acc.flag is None else int(acc.flag.item())}")
    for i, g in enumerate(gs):
        H = g.H
        rel = float(torch.linalg.norm(H - Href) / torch.linalg.norm(Href))
        print(f"   H rel err member {i}: {rel:.3e} finite={bool(torch.isfinite(H).all())}")
        g._H = None
    for i, g in enumerate(gs):
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                g.fasterquant(blocksize=128, percdamp=.01, groupsize=128, actorder=True, static_groups=True)
            print(f"   fasterquant member {i}: ok error={g.error:.4f}")
        except Exception as e:
            print(f"   fasterquant member {i}: {type(e).__name__}: {str(e)[:150]}")
        g.free()

run(False)
run(True)
run(True, members=1)
