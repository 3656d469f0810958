"""Multi-GPU partitioning of the GPTQ hot path (one process per GPU, torch.distributed).

The path shards along two axes with exactly one exchange each (BASELINE north_star):

1. calibration samples -> partial Hessians: every rank accumulates S_g = sum x x^T over ITS samples, then ONE
   all-reduce(sum) makes H identical on every rank.  Only the lower triangle travels (K (K + 1) / 2 floats, packed by a
   kernel), and the sample count rides in the same buffer, so the exchange is one collective and no host round trip.
2. W rows -> independent column loops: given the (replicated) inverse-Cholesky factor, rows of W are independent, so rank g
   quantizes rows [r0, r1) and ONE coalesced all-gather (a single NCCL group launch: codes, scales, zeros, per-shard loss --
   1 byte per weight; the dequantised fp16 weights of the other ranks' rows are rebuilt from the codes, bit for bit)
   completes the layer.

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU tests); nothing here touches
the arithmetic.
"""
from __future__ import annotations

import contextlib
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of range(n) for `rank`; shard sizes are multiples of `align`
    (except possibly the last non-empty one) and differ by at most one `align` unit."""
    if world <= 1:
        return 0, n
    units = (n + align - 1) // align
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


def sample_shard(nsamples: int, rank: int, world: int):
    """Indices of the calibration samples rank `rank` processes (round-robin, like `j mod G`)."""
    return list(range(rank, nsamples, world))


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _prof(name):
    from . import gptq as G            # late import: gptq imports nothing from here
    return G._prof(name)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if _active(group):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_hessian(gptq, group=None, total_samples: Optional[int] = None) -> None:
    """Exchange #1: sum the partial accumulators and sample counts of all ranks (in place).  Linears that share one
    accumulator (same activations: q/k/v, gate/up) are reduced once; the later ones only pick up the global count.

    `total_samples`: the global number of calibration samples when the caller knows it (the drivers do: it is the length
    of the calibration set) -- then nothing is read back from the device.  Otherwise the reduced count is read (one sync)."""
    acc = getattr(gptq, "_acc", None)
    if acc is not None and acc.reduced:
        gptq.nsamples = acc.count
        return
    S = gptq.raw_accumulator()
    K = S.shape[0]
    if _active(group):
        with _prof("nccl_allreduce"):
            if S.is_cuda:
                from . import _lib
                lib = _lib.load()
                n = K * (K + 1) // 2
                packed = torch.empty(n + 1, device=S.device, dtype=torch.float32)
                with torch.cuda.device(S.device):
                    _lib.check(lib.gptq_b200_tril_pack(S.data_ptr(), K, S.stride(0), packed.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "tril_pack")
                    packed[n:].fill_(float(gptq.nsamples))          # exact in fp32 up to 2^24 samples
                    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
                    _lib.check(lib.gptq_b200_tril_unpack(packed.data_ptr(), K, S.data_ptr(), S.stride(0),
                                                         torch.cuda.current_stream().cuda_stream), "tril_unpack")
                count = packed[n:]
            else:                                                    # CPU tensors (gloo tests of the host logic)
                idx = torch.tril_indices(K, K)
                packed = torch.cat([S[idx[0], idx[1]], torch.tensor([float(gptq.nsamples)])])
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
                S[idx[0], idx[1]] = packed[:-1]
                count = packed[-1:]
        gptq.nsamples = int(total_samples) if total_samples is not None else int(round(float(count.item())))
    acc = getattr(gptq, "_acc", None)
    if acc is not None:
        acc.count = gptq.nsamples
        acc.reduced = True


def _coalescing(group, device):
    """One NCCL group launch for several all-gathers when the backend supports it (NCCL); plain sequence otherwise."""
    try:
        from torch.distributed.distributed_c10d import _coalescing_manager
        if device.type == "cuda":
            return _coalescing_manager(group=group, device=device, async_ops=False)
    except Exception:
        pass
    return contextlib.nullcontext()


def allgather_rows_multi(locals_, rows_total: int, group=None):
    """Exchange #2 for several row-sharded tensors at once (shards from shard_range(rows_total, rank, world))."""
    if not _active(group):
        return list(locals_)
    world = dist.get_world_size(group)
    sizes = [shard_range(rows_total, r, world)[1] - shard_range(rows_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    even = all(s == mx for s in sizes)
    sends, outs = [], []
    for t in locals_:
        t = t.contiguous()
        if not even:
            pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
            t = pad
        sends.append(t)
        outs.append(torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device))
    with _coalescing(group, sends[0].device):
        for o, t in zip(outs, sends):
            dist.all_gather_into_tensor(o, t, group=group)
    if even:
        return outs
    return [torch.cat([o[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0) for o in outs]


def allgather_rows(local: torch.Tensor, rows_total: int, group=None) -> torch.Tensor:
    return allgather_rows_multi([local], rows_total, group)[0]


def fasterquant_sharded(gptq, group=None, **kw):
    """Row-sharded fasterquant: every rank solves its row shard, then the shards are gathered and the
    full quantized weight is written to ``gptq.layer`` on every rank (bitwise identical everywhere).
    ``gptq.error`` becomes the sum over all shards (a device scalar read lazily, see GPTQ.error)."""
    if not _active(group):
        return gptq.fasterquant(**kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    R = gptq.rows
    lo, hi = shard_range(R, rank, world)
    gptq.fasterquant(row_range=(lo, hi), **kw)
    q = gptq.quantizer
    per_row_q = q.scale.shape[0] == hi - lo and q.scale.dim() >= 1 and q.scale.numel() % max(hi - lo, 1) == 0 and hi - lo != R
    loss = gptq.error_tensor().reshape(1, 1).expand(hi - lo, 1).contiguous() if hi > lo else gptq.error_tensor().new_zeros((0, 1))
    # Codes, scales and zeros determine the dequantised weights bit for bit (quant.py:10 + the cast of gptq.py:305), so the
    # exchange carries 1 byte per weight instead of 3: every rank rebuilds the other ranks' rows from their codes
    # (gptq_b200_dequant_codes).  The ternary grid (maxq < 0) has no such closed form here and ships its weights.
    maxq = getattr(gptq, "_maxq", None)        # python int recorded by fasterquant (q.maxq may live on the device)
    rebuild = gptq.codes.is_cuda and maxq is not None and maxq > 0 and getattr(gptq, "g_idx", None) is not None
    parts = [gptq.codes, gptq.scales, gptq.zeros, loss]
    if not rebuild:
        parts.append(gptq.Wq)
    if per_row_q:
        parts += [q.scale.reshape(hi - lo, -1), q.zero.reshape(hi - lo, -1)]
    with _prof("nccl_allgather"):
        got = allgather_rows_multi(parts, R, group)
        gptq.codes, gptq.scales, gptq.zeros, losses = got[:4]
        nxt = 4
        if rebuild:
            from . import _lib
            Wq = torch.empty((R, gptq.columns), device=gptq.codes.device, dtype=gptq.Wq.dtype)
            sc, zr = gptq.scales.contiguous(), gptq.zeros.contiguous()
            gi = gptq.g_idx.to(torch.int32).contiguous()
            with torch.cuda.device(Wq.device):
                _lib.check(_lib.load().gptq_b200_dequant_codes(gptq.codes.data_ptr(), gptq.codes.stride(0), R, gptq.columns,
                                                               sc.data_ptr(), zr.data_ptr(), sc.shape[1], gi.data_ptr(), Wq.data_ptr(),
                                                               _lib.dt(Wq), Wq.stride(0), torch.cuda.current_stream().cuda_stream),
                           "dequant_codes")
            Wq[lo:hi] = gptq.Wq                    # (own rows: identical by construction; keeps the solver's own output authoritative)
        else:
            Wq = got[nxt]
            nxt += 1
        if per_row_q:
            q.scale, q.zero = got[nxt], got[nxt + 1]
    # one loss value per shard was replicated over the shard's rows: pick the first row of every shard, sum in rank order
    firsts = [shard_range(R, r, world)[0] for r in range(world) if shard_range(R, r, world)[1] > shard_range(R, r, world)[0]]
    gptq._error_dev = losses[firsts, 0].double().sum()
    gptq._error = None
    gptq.Wq = Wq
    layer = gptq.layer
    W = Wq.t() if type(layer).__name__ == "Conv1D" else Wq
    layer.weight.data = W.reshape(layer.weight.shape).to(layer.weight.data.dtype)
