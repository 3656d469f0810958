"""Multi-GPU partitioning of the GPTQ hot path (one process per GPU, torch.distributed).

The path shards along two axes with exactly one exchange each (BASELINE north_star):

1. calibration samples -> partial Hessians: every rank accumulates S_g = sum x x^T over ITS samples,
   then ONE all-reduce(sum) of S (and of the sample count) makes H identical on every rank;
2. W rows -> independent column loops: given the (replicated) inverse-Cholesky factor, rows of W
   are independent, so rank g quantizes rows [r0, r1) and ONE all-gather of the dequantised
   weights, codes, scales and zeros completes the layer.

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in the CPU
tests); nothing here touches the arithmetic.
"""
from __future__ import annotations

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


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_hessian(gptq, group=None) -> None:
    """Exchange #1: sum the partial accumulators and sample counts of all ranks (in place).  Linears that share one
    accumulator (same activations: q/k/v, gate/up) are reduced once; the later ones only pick up the global count."""
    acc = getattr(gptq, "_acc", None)
    if acc is not None and acc.reduced:
        gptq.nsamples = acc.count
        return
    S = gptq.raw_accumulator()
    allreduce_sum_(S, group)
    n = torch.tensor([gptq.nsamples], device=S.device, dtype=torch.int64)
    allreduce_sum_(n, group)
    gptq.nsamples = int(n.item())
    acc = getattr(gptq, "_acc", None)
    if acc is not None:
        acc.count = gptq.nsamples
        acc.reduced = True


def allgather_rows(local: torch.Tensor, rows_total: int, group=None) -> torch.Tensor:
    """Exchange #2: concatenate row shards produced with shard_range(rows_total, rank, world)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(rows_total, r, world)[1] - shard_range(rows_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0)


def fasterquant_sharded(gptq, group=None, **kw):
    """Row-sharded fasterquant: every rank solves its row shard, then the shards are gathered and the
    full quantized weight is written to ``gptq.layer`` on every rank (bitwise identical everywhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return gptq.fasterquant(**kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    R = gptq.rows
    lo, hi = shard_range(R, rank, world)
    gptq.fasterquant(row_range=(lo, hi), **kw)
    Wq = allgather_rows(gptq.Wq, R, group)
    gptq.codes = allgather_rows(gptq.codes, R, group)
    gptq.scales = allgather_rows(gptq.scales, R, group)
    gptq.zeros = allgather_rows(gptq.zeros, R, group)
    err = torch.tensor([gptq.error], device=Wq.device, dtype=torch.float64)
    allreduce_sum_(err, group)
    gptq.error = float(err.item())
    gptq.Wq = Wq
    layer = gptq.layer
    W = Wq.t() if type(layer).__name__ == "Conv1D" else Wq
    layer.weight.data = W.reshape(layer.weight.shape).to(layer.weight.data.dtype)
    q = gptq.quantizer
    if q.scale.shape[0] == hi - lo and q.scale.shape[0] != R:
        q.scale = allgather_rows(q.scale.reshape(hi - lo, -1), R, group)
        q.zero = allgather_rows(q.zero.reshape(hi - lo, -1), R, group)
