"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on the 1-GPU box): sample-sharded Hessian +
row-sharded fasterquant over NCCL must reproduce the single-GPU result."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

R, K, NS, T = 320, 512, 8, 256
CFG = dict(blocksize=128, percdamp=.01, groupsize=128, actorder=True, static_groups=True)


def _inputs():
    g = torch.Generator().manual_seed(3)
    chan = torch.exp(0.5 * torch.randn(K, generator=g))
    X = [(torch.randn(1, T, K, generator=g) * chan).half() for _ in range(NS)]
    W = (torch.randn(R, K, generator=g) * 0.02).half()
    return X, W


def _quantize(dev, X, W, world, rank):
    import torch.nn as nn
    from gptq_b200.gptq import GPTQ
    from gptq_b200.quant import Quantizer
    from gptq_b200 import dist as gd
    lin = nn.Linear(K, R, bias=False).half().to(dev)
    lin.weight.data = W.to(dev).clone()
    g = GPTQ(lin)
    g.quantizer = Quantizer()
    g.quantizer.configure(4, perchannel=True, sym=False, mse=False)
    for j in gd.sample_shard(NS, rank, world):
        g.add_batch(X[j].to(dev), None)
    if world > 1:
        gd.allreduce_hessian(g)
        gd.fasterquant_sharded(g, **CFG)
    else:
        g.fasterquant(**CFG)
    return lin.weight.data.cpu(), g.codes.cpu(), g.scales.cpu(), g.error


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        X, W = _inputs()
        Wq, codes, scales, err = _quantize(dev, X, W, world, rank)
        torch.save(dict(Wq=Wq, codes=codes, scales=scales, err=err), f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["codes"], r1["codes"]) and torch.equal(r0["Wq"], r1["Wq"])        # replicas bitwise identical
    X, W = _inputs()
    Wq, codes, scales, err = _quantize(torch.device("cuda:0"), X, W, 1, 0)
    agree = float((codes == r0["codes"]).double().mean())
    print(f"  [metric] two_gpu code agreement {agree:.6f}, err {err:.4f} vs {r0['err']:.4f}")
    assert agree >= 0.999
    assert abs(err - r0["err"]) <= 1e-3 * abs(err)
