"""CPU oracle for the GPTQ layer-wise quantization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gptq_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU arm.

It restates, function by function, what the reference (JerryZhou54/gptq) computes on
this path, using the same third-party arithmetic the reference uses (PyTorch CPU ops:
``matmul``, ``linalg.cholesky``, ``cholesky_inverse``, ``argsort``, ``round`` ...;
torch 2.11.0 here, the reference says "tested on v1.10.1").  Every function cites the
reference file:line it follows.  Parity is PINNED: ``oracle/make_golden.py`` imports the
reference's own ``gptq.py`` / ``quant.py`` from ``/root/reference`` in the build
container and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement bit-for-bit against those vectors.

Format-level exceptions (stated in DESIGN.md): the reference defines no 2-bit packed
layout (``pack2`` is by analogy with the 4-bit one -> "parity unpinned" for 2-bit
packing only) and its ``Quant3Linear.pack`` only accepts per-row ``[R,1]`` scales.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# quant.py:6-10  -- the affine grid
# --------------------------------------------------------------------------------------
def quantize(x, scale, zero, maxq):
    """quant.py:6-10.  ``maxq < 0`` selects the ternary ("trits") grid."""
    if maxq < 0:
        return (x > scale / 2).float() * scale + (x < zero / 2).float() * zero
    q = torch.clamp(torch.round(x / scale) + zero, 0, maxq)
    return scale * (q - zero)


def quantize_codes(x, scale, zero, maxq):
    """Integer codes implied by quant.py:9 (the reference never materialises them)."""
    return torch.clamp(torch.round(x / scale) + zero, 0, maxq)


@dataclass
class QParams:
    """State of a reference ``Quantizer`` (quant.py:12-35) that the path reads."""
    bits: int
    perchannel: bool = True
    sym: bool = False
    mse: bool = False
    norm: float = 2.4
    grid: int = 100
    maxshrink: float = 0.8
    trits: bool = False
    scale: Optional[torch.Tensor] = None
    zero: Optional[torch.Tensor] = None

    @property
    def maxq(self) -> int:
        # quant.py:27,34-35
        return -1 if self.trits else 2 ** self.bits - 1

    def ready(self) -> bool:
        # quant.py:130-131
        return self.scale is not None and bool(torch.all(self.scale != 0))

    def clone(self) -> "QParams":
        return QParams(self.bits, self.perchannel, self.sym, self.mse, self.norm, self.grid,
                       self.maxshrink, self.trits,
                       None if self.scale is None else self.scale.clone(),
                       None if self.zero is None else self.zero.clone())


def find_params(x: torch.Tensor, qp: QParams, weight: bool = True):
    """quant.py:37-118.  Returns (scale, zero) shaped like the reference leaves them."""
    maxq = qp.maxq
    shape = x.shape
    if qp.perchannel:
        if weight:
            x = x.flatten(1)
        else:
            if len(shape) == 4:
                x = x.permute([1, 0, 2, 3]).flatten(1)
            if len(shape) == 3:
                x = x.reshape((-1, shape[-1])).t()
            if len(shape) == 2:
                x = x.t()
    else:
        x = x.flatten().unsqueeze(0)

    tmp = torch.zeros(x.shape[0])
    xmin = torch.minimum(x.min(1)[0], tmp)          # quant.py:57
    xmax = torch.maximum(x.max(1)[0], tmp)          # quant.py:58
    if qp.sym:                                      # quant.py:60-64
        xmax = torch.maximum(torch.abs(xmin), xmax)
        neg = xmin < 0
        if torch.any(neg):
            xmin[neg] = -xmax[neg]
    flat = (xmin == 0) & (xmax == 0)                # quant.py:65-67
    xmin[flat] = -1
    xmax[flat] = +1

    if maxq < 0:                                    # quant.py:69-71
        scale = xmax
        zero = xmin
    else:                                           # quant.py:72-77
        scale = (xmax - xmin) / maxq
        if qp.sym:
            zero = torch.full_like(scale, (maxq + 1) / 2)
        else:
            zero = torch.round(-xmin / scale)

    if qp.mse:                                      # quant.py:79-96
        best = torch.full([x.shape[0]], float("inf"))
        for i in range(int(qp.maxshrink * qp.grid)):
            p = 1 - i / qp.grid
            xmin1 = p * xmin
            xmax1 = p * xmax
            scale1 = (xmax1 - xmin1) / maxq
            zero1 = torch.round(-xmin1 / scale1) if not qp.sym else zero
            q = quantize(x, scale1.unsqueeze(1), zero1.unsqueeze(1), maxq)
            q -= x
            q.abs_()
            q.pow_(qp.norm)
            err = torch.sum(q, 1)
            better = err < best
            if torch.any(better):
                best[better] = err[better]
                scale[better] = scale1[better]
                zero[better] = zero1[better]
    if not qp.perchannel:                           # quant.py:97-103
        rep = shape[0] if weight else (shape[1] if len(shape) != 3 else shape[2])
        scale = scale.repeat(rep)
        zero = zero.repeat(rep)
    if weight:                                      # quant.py:105-109
        shp = [-1] + [1] * (len(shape) - 1)
        return scale.reshape(shp), zero.reshape(shp)
    if len(shape) == 4:
        return scale.reshape((1, -1, 1, 1)), zero.reshape((1, -1, 1, 1))
    if len(shape) == 3:
        return scale.reshape((1, 1, -1)), zero.reshape((1, 1, -1))
    return scale.unsqueeze(0), zero.unsqueeze(0)


# --------------------------------------------------------------------------------------
# gptq.py:25-65  -- Hessian accumulation
# --------------------------------------------------------------------------------------
@dataclass
class HessianState:
    columns: int
    H: torch.Tensor = None
    nsamples: int = 0
    last_input_mean: Optional[torch.Tensor] = None

    def __post_init__(self):
        if self.H is None:
            self.H = torch.zeros((self.columns, self.columns))   # gptq.py:35


def add_batch(st: HessianState, inp: torch.Tensor) -> None:
    """gptq.py:38-65 for nn.Linear / Conv1D layers (running-mean Hessian)."""
    if len(inp.shape) == 2:
        inp = inp.unsqueeze(0)
    tmp = inp.shape[0]
    if len(inp.shape) == 3:
        inp = inp.reshape((-1, inp.shape[-1]))
    inp = inp.t()
    st.H *= st.nsamples / (st.nsamples + tmp)        # gptq.py:59
    st.nsamples += tmp
    inp = math.sqrt(2 / st.nsamples) * inp.float()    # gptq.py:62
    st.last_input_mean = torch.mean(inp, 1)           # gptq.py:63
    st.H += inp.matmul(inp.t())                       # gptq.py:65


def hessian_closed_form(batches: List[torch.Tensor]) -> torch.Tensor:
    """H = (2/N) * sum_j X_j^T X_j in fp64 -- what gptq.py:59-65 converges to."""
    K = batches[0].shape[-1]
    H = torch.zeros((K, K), dtype=torch.float64)
    for b in batches:
        x = b.reshape(-1, K).double()
        H += x.t() @ x
    return H * (2.0 / len(batches))


# --------------------------------------------------------------------------------------
# gptq.py:143-145, 165-180 -- dead columns, act-order, damp + 3-step Cholesky
# --------------------------------------------------------------------------------------
def hinv_upper(H: torch.Tensor, percdamp: float) -> torch.Tensor:
    """gptq.py:174-180.  H is modified in place like the reference does."""
    K = H.shape[0]
    damp = percdamp * torch.mean(torch.diag(H))
    diag = torch.arange(K)
    H[diag, diag] += damp
    H = torch.linalg.cholesky(H)
    H = torch.cholesky_inverse(H)
    H = torch.linalg.cholesky(H, upper=True)
    return H


@dataclass
class FasterquantResult:
    Q: torch.Tensor                 # [R,K] fp32 dequantised weights, original column order (gptq.py:300-305)
    codes: torch.Tensor             # [R,K] uint8 integer codes, original column order (implied by quant.py:9)
    scales: torch.Tensor            # [R, ngroups] scale actually used for each group
    zeros: torch.Tensor             # [R, ngroups]
    g_idx: torch.Tensor             # [K] group index of each ORIGINAL column
    perm: Optional[torch.Tensor]    # act-order permutation (gptq.py:166) or None
    error: float                    # sum(Losses)  (gptq.py:294)
    final_scale: torch.Tensor       # quantizer.scale after the run ([R,1]; last group)
    final_zero: torch.Tensor
    U: torch.Tensor = field(repr=False, default=None)   # upper factor of H^-1 (gptq.py:179)
    margin: torch.Tensor = field(repr=False, default=None)   # [R,K] distance of w/scale from the nearest rounding
                                                             # boundary (x.5) when the column was quantized (keep_margin)


def fasterquant(W: torch.Tensor, H: torch.Tensor, qp: QParams, blocksize: int = 128,
                percdamp: float = .01, groupsize: int = -1, actorder: bool = False,
                static_groups: bool = False, keep_U: bool = False, keep_margin: bool = False,
                chain_fp64: bool = False, stable_perm: bool = False) -> FasterquantResult:
    """gptq.py:126-305, standard branch (lut_quant = non_linear_quant = columnwise = False).

    ``W`` [R,K] (any float dtype), ``H`` [K,K] fp32 as left by add_batch.  Both are cloned.
    """
    W = W.clone().float()                                   # gptq.py:130-135
    H = H.clone()
    R, K = W.shape
    maxq = qp.maxq
    qp = qp.clone()

    dead = torch.diag(H) == 0                               # gptq.py:143-145
    H[dead, dead] = 1
    W[:, dead] = 0

    groups = []
    if static_groups:                                       # gptq.py:157-163
        for i in range(0, K, groupsize):
            g = qp.clone()
            g.scale, g.zero = find_params(W[:, i:(i + groupsize)], g)
            groups.append(g)

    perm = None
    if actorder:                                            # gptq.py:165-169
        # gptq.py:166 does not ask for a stable sort: EXACT ties in diag(H) (two channels with the same fp32 sum of squares;
        # several dead columns with diag = 1) come out in an implementation-defined order that differs between torch's CPU
        # sort (this line, what the golden vectors pin) and its CUDA radix sort (stable: ties keep ascending index, which is
        # also what gptq_b200_argsort_desc produces).  `stable_perm` selects the latter for comparisons with the CUDA path.
        perm = torch.argsort(torch.diag(H), descending=True, stable=stable_perm)
        W = W[:, perm]
        H = H[perm][:, perm]
        invperm = torch.argsort(perm)

    Losses = torch.zeros_like(W)
    Q = torch.zeros_like(W)
    C = torch.zeros_like(W)
    M = torch.full_like(W, 0.5) if keep_margin else None     # test instrumentation only; not part of the reference

    if chain_fp64:      # CONTROL EXPERIMENT (tests only, not the reference): the exact factor, rounded once to fp32
        Hinv = hinv_upper(H.double(), percdamp).float()
    else:
        Hinv = hinv_upper(H, percdamp)                      # gptq.py:174-180
    if not qp.ready():                                      # gptq.py:181-185
        qp.scale, qp.zero = find_params(W, qp)

    ngroups = 1 if groupsize == -1 else (K + groupsize - 1) // groupsize
    scales = torch.zeros((R, ngroups))
    zeros = torch.zeros((R, ngroups))
    g_idx_perm = torch.zeros(K, dtype=torch.int64)
    if groupsize == -1:
        scales[:, 0] = qp.scale.flatten()
        zeros[:, 0] = qp.zero.flatten()

    cur = qp
    for i1 in range(0, K, blocksize):                       # gptq.py:191-276
        i2 = min(i1 + blocksize, K)
        count = i2 - i1
        W1 = W[:, i1:i2].clone()
        Q1 = torch.zeros_like(W1)
        C1 = torch.zeros_like(W1)
        Err1 = torch.zeros_like(W1)
        Losses1 = torch.zeros_like(W1)
        Hinv1 = Hinv[i1:i2, i1:i2]
        for i in range(count):
            w = W1[:, i]
            d = Hinv1[i, i]
            if groupsize != -1:                             # gptq.py:252-260
                if not static_groups:
                    if (i1 + i) % groupsize == 0:
                        cur.scale, cur.zero = find_params(W[:, (i1 + i):(i1 + i + groupsize)], cur)
                    g = (i1 + i) // groupsize
                else:
                    idx = i1 + i
                    if actorder:
                        idx = int(perm[idx])
                    g = idx // groupsize
                    cur = groups[g]
                g_idx_perm[i1 + i] = g
                scales[:, g] = cur.scale.flatten()
                zeros[:, g] = cur.zero.flatten()
            q = quantize(w.unsqueeze(1), cur.scale, cur.zero, maxq).flatten()   # gptq.py:262-264
            if maxq > 0:
                C1[:, i] = quantize_codes(w.unsqueeze(1), cur.scale, cur.zero, maxq).flatten()
                if keep_margin:
                    t = (w.unsqueeze(1) / cur.scale).flatten()
                    M[:, i1 + i] = torch.abs(t - torch.floor(t) - 0.5)
            Q1[:, i] = q
            Losses1[:, i] = (w - q) ** 2 / d ** 2           # gptq.py:267
            err1 = (w - q) / d                              # gptq.py:269
            W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0))   # gptq.py:270
            Err1[:, i] = err1
        Q[:, i1:i2] = Q1
        C[:, i1:i2] = C1
        Losses[:, i1:i2] = Losses1 / 2                      # gptq.py:274
        W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])          # gptq.py:276

    error = torch.sum(Losses).item()                        # gptq.py:294
    if actorder:                                            # gptq.py:300-301
        Q = Q[:, invperm]
        C = C[:, invperm]
        g_idx = g_idx_perm[invperm]
        if M is not None:
            M = M[:, invperm]
    else:
        g_idx = g_idx_perm
    return FasterquantResult(Q=Q, codes=C.to(torch.uint8), scales=scales, zeros=zeros, g_idx=g_idx,
                             perm=perm, error=error, final_scale=cur.scale.clone(),
                             final_zero=cur.zero.clone(), U=Hinv if keep_U else None, margin=M)


# --------------------------------------------------------------------------------------
# Packing  (quant.py:152-187, zeroShot/models/quant.py:172-185)
# --------------------------------------------------------------------------------------
def intweight_from_linear(weight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor) -> np.ndarray:
    """quant.py:153-160: recover integer codes from the dequantised weights.

    weight [R,K] (fp16 in the drivers), scales/zeros [R,1] fp32 -> uint32 [K,R].
    """
    zs = zeros * scales                                       # quant.py:153
    intweight = torch.round((weight + zs) / scales).to(torch.int)   # quant.py:158
    intweight = intweight.t().contiguous()
    return intweight.numpy().astype(np.uint32)


def pack3(intweight: np.ndarray) -> np.ndarray:
    """quant.py:161-186.  intweight uint32 [K,R] (K % 32 == 0) -> int32 [K//32*3, R].

    Per 32 input rows: an LSB-first 96-bit stream of 32 3-bit codes in 3 uint32 words.
    Written as the stream (not as the reference's unrolled loop); tests check both agree.
    """
    K, R = intweight.shape
    assert K % 32 == 0
    iw = intweight.astype(np.uint64).reshape(K // 32, 32, R)
    out = np.zeros((K // 32, 3, R), dtype=np.uint64)
    for j in range(32):
        bit = 3 * j
        word, off = divmod(bit, 32)
        out[:, word, :] |= (iw[:, j, :] << np.uint64(off)) & np.uint64(0xFFFFFFFF)
        if off > 29:                                          # code straddles two words
            out[:, word + 1, :] |= iw[:, j, :] >> np.uint64(32 - off)
    return out.reshape(K // 32 * 3, R).astype(np.uint32).view(np.int32)


def pack3_reference_loop(intweight: np.ndarray) -> np.ndarray:
    """Literal transcription of the control flow at quant.py:164-184 (slow; small inputs)."""
    qweight = np.zeros((intweight.shape[0] // 32 * 3, intweight.shape[1]), dtype=np.uint32)
    i = 0
    row = 0
    while row < qweight.shape[0]:
        for j in range(i, i + 10):
            qweight[row] |= intweight[j] << (3 * (j - i))
        i += 10
        qweight[row] |= intweight[i] << 30
        row += 1
        qweight[row] |= (intweight[i] >> 2) & 1
        i += 1
        for j in range(i, i + 10):
            qweight[row] |= intweight[j] << (3 * (j - i) + 1)
        i += 10
        qweight[row] |= intweight[i] << 31
        row += 1
        qweight[row] |= (intweight[i] >> 1) & 0x3
        i += 1
        for j in range(i, i + 10):
            qweight[row] |= intweight[j] << (3 * (j - i) + 2)
        i += 10
        row += 1
    return qweight.astype(np.int32)


def pack_pow2(intweight: np.ndarray, bits: int) -> np.ndarray:
    """4-bit: zeroShot/models/quant.py:184-185 ``qweight[i//8] |= code[i] << 4*(i%8)``.

    2-bit (and 8-bit) by analogy: ``qweight[i//(32/b)] |= code[i] << b*(i % (32/b))``.
    The reference defines no 2-bit layout -> parity unpinned for bits=2.
    """
    per = 32 // bits
    K, R = intweight.shape
    assert K % per == 0
    iw = intweight.astype(np.uint32).reshape(K // per, per, R)
    out = np.zeros((K // per, R), dtype=np.uint32)
    for j in range(per):
        out |= iw[:, j, :] << np.uint32(bits * j)
    return out.view(np.int32)


def pack(intweight: np.ndarray, bits: int) -> np.ndarray:
    if bits == 3:
        return pack3(intweight)
    return pack_pow2(intweight, bits)


def unpack(qweight: np.ndarray, bits: int, K: int) -> np.ndarray:
    """Inverse of ``pack`` (mirrors the unpack order of quant_cuda_kernel.cu:116-162)."""
    qw = qweight.view(np.uint32).astype(np.uint64)
    R = qw.shape[1]
    if bits != 3:
        per = 32 // bits
        out = np.zeros((K, R), dtype=np.uint32)
        for j in range(per):
            out[j::per] = ((qw >> np.uint64(bits * j)) & np.uint64((1 << bits) - 1)).astype(np.uint32)
        return out
    qw = qw.reshape(K // 32, 3, R)
    out = np.zeros((K // 32, 32, R), dtype=np.uint32)
    for j in range(32):
        word, off = divmod(3 * j, 32)
        v = qw[:, word, :] >> np.uint64(off)
        if off > 29:
            v |= qw[:, word + 1, :] << np.uint64(32 - off)
        out[:, j, :] = (v & np.uint64(7)).astype(np.uint32)
    return out.reshape(K, R)


def vecquant_matvec(x: np.ndarray, qweight: np.ndarray, scales: np.ndarray, zeros_scaled: np.ndarray, bits: int,
                    bias: Optional[np.ndarray] = None, g_idx: Optional[np.ndarray] = None) -> np.ndarray:
    """quant_cuda_kernel.cu:88-165 in float64: y[c] = bias[c] + sum_k (scale[c,g]*code[k,c] - zs[c,g]) * x[k]."""
    K = x.shape[0]
    codes = unpack(qweight, bits, K).astype(np.float64)          # [K, R]
    R = codes.shape[1]
    s = scales.reshape(R, -1).astype(np.float64)
    z = zeros_scaled.reshape(R, -1).astype(np.float64)
    g = np.zeros(K, dtype=np.int64) if g_idx is None else np.asarray(g_idx, dtype=np.int64)
    Wd = s[:, g].T * codes - z[:, g].T                           # [K, R]
    y = x.astype(np.float64) @ Wd
    return y if bias is None else y + bias.astype(np.float64)


# --------------------------------------------------------------------------------------
# Parity metrics
# --------------------------------------------------------------------------------------
def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double()
    b = b.double()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def recon_error(W: torch.Tensor, Q: torch.Tensor, H: torch.Tensor) -> float:
    """||W X - Q X||^2 (up to the 2/N factor in H) = tr(dW H dW^T)."""
    d = (W.double() - Q.double())
    return float(torch.einsum("ik,kl,il->", d, H.double(), d))


EXACT_TIE = 2.0 ** -20        # |frac(w/scale) - 0.5| below this: the tie is at the resolution of fp32 itself


def classify_mismatches(codes: torch.Tensor, ref: "FasterquantResult", tol: float, perm: Optional[torch.Tensor] = None,
                        diag: Optional[torch.Tensor] = None, perm_tol: float = 1e-5) -> dict:
    """Rounding-tie classifier for a run compared with the oracle run `ref` (which must carry `margin`).

    Rows are independent recurrences (gptq.py:191-276 acts row-wise), and inside a row everything after the first
    differing column legitimately differs (the error feedback of that column is different).  So a row's mismatches are
    explained iff its FIRST differing column, in processing order (act-order permutation), was
      * an exact rounding tie for the oracle: |frac(w/scale) - 0.5| < EXACT_TIE = 2^-20 (not reproducible by ANY fp32
        implementation that reorders a sum), or
      * a near tie: margin < `tol`, i.e. a perturbation of w/scale below `tol` quantization steps flips round(), or
      * (act-order, when the other run's `perm` and diag(H) are given) at or after a SORT-KEY tie: the two runs process
        different columns at some earlier-or-equal position and the diag(H) values being permuted there agree to `perm_tol`
        relative -- the act-order argsort of gptq.py:166 is discontinuous at such ties, so Hessians equal to 1e-5 legitimately
        give different processing orders, and from that position on the two runs are different (equally valid) GPTQ
        solutions.  Positions where the orders differ WITHOUT such a tie are reported as `perm_unexplained_positions`.
    `mismatch_outside_exact_tie_rows` is the code mismatch over the whole matrix when the tails of exact-tie rows are not counted.
    """
    assert ref.margin is not None, "run the oracle with keep_margin=True"
    a, b, m = codes.cpu(), ref.codes, ref.margin
    swapped = torch.zeros(a.shape[1], dtype=torch.bool)
    n_swapped = n_perm_bad = 0
    if ref.perm is not None:                      # original order -> processing order
        a, b, m = a[:, ref.perm], b[:, ref.perm], m[:, ref.perm]
        if perm is not None:
            pg = perm.cpu().long()
            differs = pg != ref.perm
            n_swapped = int(differs.sum())
            if n_swapped and diag is not None:
                d = diag.cpu().double()
                close = (d[pg] - d[ref.perm]).abs() <= perm_tol * d[ref.perm].abs()
                n_perm_bad = int((differs & ~close).sum())
                if bool((differs & close).any()):
                    swapped[int(torch.nonzero(differs & close)[0]):] = True       # everything from the first sort-key tie on
    diff = a != b
    bad_rows = torch.nonzero(diff.any(dim=1)).flatten()
    first = diff[bad_rows].to(torch.uint8).argmax(dim=1)
    fm = m[bad_rows, first] if len(bad_rows) else torch.zeros(0)
    exact = fm < EXACT_TIE
    tie = fm < tol
    at_swap = swapped[first] if len(bad_rows) else torch.zeros(0, dtype=torch.bool)
    explained = tie | at_swap
    unexplained = bad_rows[~explained]
    per_row = diff[bad_rows].sum(dim=1) if len(bad_rows) else torch.zeros(0, dtype=torch.int64)
    unexp_m = fm[~explained]
    return dict(mismatch=float(diff.double().mean()), rows_bad=int(len(bad_rows)), exact_tie_rows=int(exact.sum()),
                tie_rows=int(tie.sum()), sort_key_tie_rows=int((at_swap & ~tie).sum()), perm_positions_differing=n_swapped,
                perm_unexplained_positions=n_perm_bad,
                unexplained_rows=int(len(unexplained)),
                mismatch_outside_exact_tie_rows=float(per_row[~exact].sum()) / diff.numel() if len(bad_rows) else 0.0,
                max_rounding_margin=float(fm[~at_swap | tie].max()) if len(fm) and bool((~at_swap | tie).any()) else 0.0,
                max_unexplained_margin=float(unexp_m.max()) if len(unexp_m) else 0.0,
                first_margins=[float(x) for x in sorted(fm.tolist(), reverse=True)[:6]],
                first_cols=[int(first[i]) for i in (torch.argsort(fm, descending=True)[:6].tolist() if len(fm) else [])], tol=tol)


def code_agreement(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a == b).double().mean())
