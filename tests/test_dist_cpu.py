"""Host-side multi-GPU logic on CPU: partition arithmetic and the two exchanges (gloo, world size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gptq_b200 import dist as gd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 768, 4096, 11008, 12288):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 16):
                spans = [gd.shard_range(n, r, world, align) for r in range(world)]
                assert spans[0][0] == 0 and spans[-1][1] == n
                for (a, b), (c, d) in zip(spans, spans[1:]):
                    assert b == c and a <= b
                sizes = [b - a for a, b in spans]
                assert max(sizes) - min(sizes) <= align
                if align > 1:
                    assert all(a % align == 0 for a, b in spans if b > a)


def test_sample_shard_covers_all_samples():
    for world in (1, 2, 4, 8):
        seen = sorted(j for r in range(world) for j in gd.sample_shard(128, r, world))
        assert seen == list(range(128))


class _FakeGPTQ:
    def __init__(self, S, n):
        self._S, self.nsamples = S, n

    def raw_accumulator(self):
        return self._S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # exchange #1: partial Hessians + sample counts
        g = torch.Generator().manual_seed(rank)
        X = torch.randn(64, 16, generator=g)
        fake = _FakeGPTQ(X.t() @ X, 64 // 2 + rank)
        mine = (X.t() @ X).clone()
        gd.allreduce_hessian(fake)
        want = sum((torch.randn(64, 16, generator=torch.Generator().manual_seed(r)).t()
                    @ torch.randn(64, 16, generator=torch.Generator().manual_seed(r))) for r in range(world))
        # only the lower triangle travels (the consumers mirror it): lower = global sum, strict upper = untouched
        assert torch.allclose(torch.tril(fake._S), torch.tril(want), atol=1e-5)
        assert torch.equal(torch.triu(fake._S, 1), torch.triu(mine, 1))
        assert fake.nsamples == sum(64 // 2 + r for r in range(world))
        # a caller that knows the global sample count passes it: nothing is read back
        fake2 = _FakeGPTQ(mine.clone(), 3)
        gd.allreduce_hessian(fake2, total_samples=77)
        assert fake2.nsamples == 77 and torch.allclose(torch.tril(fake2._S), torch.tril(want), atol=1e-5)
        # linears sharing one accumulator (same activations) are reduced ONCE; the siblings only pick up the global count
        class _Acc:
            reduced, count = False, 0
        acc = _Acc()
        S = X.t() @ X
        a, b = _FakeGPTQ(S, 10 + rank), _FakeGPTQ(S, 10 + rank)
        a._acc = b._acc = acc
        gd.allreduce_hessian(a)
        gd.allreduce_hessian(b)
        assert torch.allclose(torch.tril(S), torch.tril(want), atol=1e-5)   # summed once, not twice
        assert a.nsamples == b.nsamples == acc.count == sum(10 + r for r in range(world)) and acc.reduced
        # exchange #2: row shards (uneven on purpose)
        full = torch.arange(rows * 5, dtype=torch.float32).reshape(rows, 5)
        lo, hi = gd.shard_range(rows, rank, world)
        got = gd.allgather_rows(full[lo:hi].clone(), rows)
        assert torch.equal(got, full)
        codes = (torch.arange(rows * 8) % 251).to(torch.uint8).reshape(rows, 8)
        assert torch.equal(gd.allgather_rows(codes[lo:hi].clone(), rows), codes)
        # the row-sharded solve: every rank quantizes its rows, one exchange completes the layer on every rank
        import torch.nn as nn
        lin = nn.Linear(5, rows, bias=False)
        Wq_full = (full / 7).half()
        scales = torch.arange(rows * 2, dtype=torch.float32).reshape(rows, 2) + 1

        class _Q:
            scale = torch.zeros(1)
            zero = torch.zeros(1)

        class _Sharded:
            def __init__(self):
                self.layer, self.rows, self.quantizer = lin, rows, _Q()

            def fasterquant(self, row_range=None, **kw):
                a, b = row_range
                self.Wq, self.codes = Wq_full[a:b].clone(), codes[a:b, :5].clone()
                self.scales, self.zeros = scales[a:b].clone(), -scales[a:b].clone()
                self._e = torch.tensor([float(b - a)], dtype=torch.float64)
                self.quantizer.scale = scales[a:b, :1].clone()
                self.quantizer.zero = scales[a:b, 1:].clone()

            def error_tensor(self):
                return self._e
        sh = _Sharded()
        gd.fasterquant_sharded(sh, blocksize=128)
        assert torch.equal(sh.Wq, Wq_full) and torch.equal(lin.weight.data, Wq_full.float())
        assert torch.equal(sh.codes, codes[:, :5]) and torch.equal(sh.scales, scales) and torch.equal(sh.zeros, -scales)
        assert float(sh._error_dev) == float(rows)                     # sum of the shards' losses, each counted once
        assert torch.equal(sh.quantizer.scale, scales[:, :1]) and torch.equal(sh.quantizer.zero, scales[:, 1:])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rows", [8, 11])
def test_exchanges_world2_gloo(rows):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, rows), nprocs=2, join=True)
