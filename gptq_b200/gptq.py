"""Host-side mirror of the reference's ``gptq.py`` (JerryZhou54/gptq) over the sm_100a C ABI.

``GPTQ(layer).add_batch(inp, out)`` / ``fasterquant(...)`` / ``free()`` keep the reference's
signatures, attributes (``layer, dev, rows, columns, H, nsamples, quantizer, input``) and side
effects (``fasterquant`` consumes ``H``, overwrites ``layer.weight.data`` in place, prints
``time`` / ``error``, raises ``torch.linalg.LinAlgError`` for a non-PD Hessian) -- gptq.py:23-318.

What differs underneath:

* ``add_batch`` stages the activations (fp16/bf16, as produced by the model) in a device slab and
  folds ``STAGE_SAMPLES`` samples at a time into the unscaled accumulator S = sum x x^T with the
  tcgen05 syrk (lower tiles only).  ``H`` = (2/nsamples) * S, mirrored, is materialised when
  read -- the closed form of the running update at gptq.py:59-65.
* ``fasterquant`` runs the prologue, the inverse-Cholesky chain, the column loop and the
  un-permute/cast epilogue as C-ABI calls on the current CUDA stream; the only host syncs are the
  ones the reference also has (``quantizer.ready()``, the final ``synchronize`` + ``.item()``).

Extras the reference does not expose (needed for packing grouped models, tests and sharding):
``codes`` (uint8 [R,K]), ``scales`` / ``zeros`` ([R,ngroups]), ``g_idx`` ([K]), ``perm``, ``error``.
"""
from __future__ import annotations

import copy
import math
import os
import time

import torch
import torch.nn as nn

try:  # transformers is only needed to recognise Conv1D (GPT-2 style) layers, like the reference
    import transformers
    _Conv1D = transformers.Conv1D
except Exception:  # pragma: no cover
    transformers = None

    class _Conv1D:  # type: ignore
        pass

from . import _lib
from .quant import Quantizer, find_params_groups

DEBUG = False

torch.backends.cuda.matmul.allow_tf32 = False     # gptq.py:18-19
torch.backends.cudnn.allow_tf32 = False

STAGE_SAMPLES = int(os.environ.get("GPTQ_B200_STAGE_SAMPLES", "16"))
CHUNK_TOKENS = int(os.environ.get("GPTQ_B200_CHUNK_TOKENS", "0"))     # 0 = library default
EMPTY_CACHE_ON_FREE = os.environ.get("GPTQ_B200_EMPTY_CACHE", "0") == "1"
CHAIN_MEMO = os.environ.get("GPTQ_B200_CHAIN_MEMO", "1") == "1"
_CHAIN_MEMO = {}          # (device, Kp) -> (A as handed to the chain, U, info); at most one entry


# Optional per-call device timing: set ``gptq_b200.gptq.PROFILE = []`` and every C-ABI hot call appends
# (name, start_event, end_event, work) recorded on the launching stream (bench.py reads it after a sync).
PROFILE = None


class _prof:
    def __init__(self, name, work=0.0):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.append((self.name, self.a, self.b, self.work))
        return False


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def syrk_issued_flops(K: int, tokens: int) -> float:
    """MMA flops the syrk kernel actually issues: 128x256 tiles touching the lower triangle."""
    nbm, nbn = (K + 127) // 128, (K + 255) // 256
    tiles = sum(nbm - 2 * j for j in range(nbn))
    return 2.0 * tiles * 128 * 256 * tokens


class GPTQ:

    def __init__(self, layer):
        _CHAIN_MEMO.clear()            # a new calibration round starts: drop the previous round's chain memo
        self.layer = layer
        self.dev = self.layer.weight.device
        _lib.require_cuda(layer.weight, "GPTQ: layer.weight")
        _lib.load()
        W = layer.weight.data
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        self.rows = W.shape[0]
        self.columns = W.shape[1]
        self.nsamples = 0
        self.quantizer = None
        # unscaled accumulator S = sum_tokens x x^T (fp32, tiles touching the lower triangle valid)
        self._S = torch.zeros((self.columns, self.columns), device=self.dev, dtype=torch.float32)
        self._H = None                 # materialised (2/n)*S, full symmetric
        self._stage = None             # [cap_tokens, ld] staging slab
        self._staged = 0
        self._last = None              # (view of the last batch in the slab, nsamples at that time)
        self._ws = None

    # ------------------------------------------------------------------ Hessian (gptq.py:38-65)
    def add_batch(self, inp, out):
        if DEBUG:
            self.inp1 = inp
            self.out1 = out
        _lib.require_cuda(inp, "GPTQ.add_batch: inp")
        if len(inp.shape) == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        if isinstance(self.layer, nn.Linear) or isinstance(self.layer, _Conv1D):
            if len(inp.shape) == 3:
                inp = inp.reshape((-1, inp.shape[-1]))            # [tokens, K]; the reference transposes, we don't need to
        if isinstance(self.layer, nn.Conv2d):
            unfold = nn.Unfold(self.layer.kernel_size, dilation=self.layer.dilation, padding=self.layer.padding,
                               stride=self.layer.stride)
            inp = unfold(inp).permute([1, 0, 2]).flatten(1).t()   # [tokens, K]
        if inp.shape[-1] != self.columns:
            raise ValueError(f"add_batch: expected {self.columns} input features, got {inp.shape[-1]}")
        self.nsamples += tmp
        self._H = None
        if inp.dtype == torch.float32:
            # fp32 activations: exact fp32 CUDA-core path (the tensor path takes 16-bit operands)
            self._flush()
            x = inp.contiguous()
            _lib.check(_lib.load().gptq_b200_syrk_accum_simt(x.data_ptr(), _lib.F32, x.shape[0], self.columns, x.stride(0),
                                                             self._S.data_ptr(), self.columns, 1.0, 1.0, _lib.stream()),
                       "syrk_accum_simt")
            self._last = (x, self.nsamples)
            return
        self._stage_tokens(inp)

    def _stage_tokens(self, x):
        lib = _lib.load()
        K = self.columns
        tokens = x.shape[0]
        if x.stride(-1) != 1:
            x = x.contiguous()
        if self._stage is None or self._stage.dtype != x.dtype:
            self._flush()
            ld = _round_up(K, 8)
            cap = max(tokens, 1) * max(STAGE_SAMPLES, 1)
            self._stage = torch.empty((cap, ld), device=self.dev, dtype=x.dtype)
        cap = self._stage.shape[0]
        done = 0
        while done < tokens:
            if self._staged == cap:
                self._flush()
            n = min(tokens - done, cap - self._staged)
            dst = self._stage[self._staged:self._staged + n]
            src = x[done:done + n]
            _lib.check(lib.gptq_b200_stage_input(src.data_ptr(), _lib.dt(src), n, K, src.stride(0), dst.data_ptr(),
                                                 dst.stride(0), _lib.stream()), "stage_input")
            self._last = (dst[:, :K], self.nsamples)
            self._staged += n
            done += n

    def _flush(self):
        if self._stage is None or self._staged == 0:
            return
        lib = _lib.load()
        K = self.columns
        need = lib.gptq_b200_syrk_workspace_bytes(K, self._staged)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        with _prof("syrk", syrk_issued_flops(K, self._staged)):
            _lib.check(lib.gptq_b200_syrk_accum(self._stage.data_ptr(), _lib.dt(self._stage), self._staged, K,
                                                self._stage.stride(0), self._S.data_ptr(), K, 1.0, 1.0, CHUNK_TOKENS,
                                                _lib.ptr(self._ws) if need else 0, need, _lib.stream()), "syrk_accum")
        self._staged = 0

    @property
    def H(self):
        """(2/nsamples) * sum x x^T as a full symmetric fp32 matrix (what gptq.py:59-65 maintains)."""
        if self._H is None and self._S is not None:
            self._flush()
            H = self._S.clone()
            _lib.check(_lib.load().gptq_b200_symmetrize(H.data_ptr(), self.columns, self.columns, _lib.stream()),
                       "symmetrize")
            if self.nsamples > 0:
                H.mul_(2.0 / self.nsamples)
            self._H = H
        return self._H

    @H.setter
    def H(self, value):
        self._H = value
        if value is None:
            self._S = None

    @H.deleter
    def H(self):
        self._H = None

    @property
    def input(self):
        """gptq.py:63: mean over tokens of sqrt(2/n) * (last batch)."""
        if self._last is None:
            raise AttributeError("input")
        x, n = self._last
        return math.sqrt(2 / n) * x.float().mean(0)

    def raw_accumulator(self):
        """Unscaled S (after folding in anything still staged) -- what multi-GPU runs all-reduce."""
        self._flush()
        self._H = None
        return self._S

    # ------------------------------------------------------------------ solve (gptq.py:126-309)
    def fasterquant(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False,
                    model_name="opt", layer_name="layer", lut_quant=False, non_linear_quant=False, columnwise=False,
                    row_range=None):
        if lut_quant or non_linear_quant or columnwise:
            raise NotImplementedError("lut_quant / non_linear_quant / columnwise are research modes of the fork "
                                      "outside the GPTQ hot path (SURVEY 8a)")
        if not isinstance(self.quantizer, Quantizer):
            raise TypeError("GPTQ.quantizer must be a gptq_b200.quant.Quantizer")
        lib = _lib.load()
        st = _lib.stream()
        dev = self.dev
        q = self.quantizer
        K = self.columns
        W = self.layer.weight.data
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        if row_range is not None:                       # row sharding: rows are independent given H
            W = W[row_range[0]:row_range[1]]
        W = W.contiguous()
        R = W.shape[0]
        maxq = int(q.maxq.item())
        if maxq == 0:
            raise RuntimeError("Quantizer.configure() was not called")

        tick = time.time()
        if self._H is None and self._S is not None:     # consume the accumulator in place (no clone)
            self._flush()
            H = self._S
            _lib.check(lib.gptq_b200_symmetrize(H.data_ptr(), K, K, st), "symmetrize")
            if self.nsamples > 0:
                H.mul_(2.0 / self.nsamples)
        else:
            H = self._H                                 # a user-provided / already materialised H (gptq.py:141-142)
        self._H = None
        self._S = None
        self._stage = None
        self._last = None
        if H is None:
            raise RuntimeError("fasterquant: H was already consumed or freed")

        f32 = dict(device=dev, dtype=torch.float32)
        _p0 = _prof("prologue")
        _p0.__enter__()
        dead = torch.empty(K, device=dev, dtype=torch.uint8)
        diag = torch.empty(K, **f32)
        damp = torch.empty(1, **f32)
        ndead = torch.empty(1, device=dev, dtype=torch.int32)
        _lib.check(lib.gptq_b200_dead_columns(H.data_ptr(), K, K, dead.data_ptr(), st), "dead_columns")
        _lib.check(lib.gptq_b200_hessian_prepare(H.data_ptr(), K, K, float(percdamp), diag.data_ptr(), damp.data_ptr(),
                                                 ndead.data_ptr(), st), "hessian_prepare")
        W32 = torch.empty((R, K), **f32)
        _lib.check(lib.gptq_b200_weights_prepare(W.data_ptr(), _lib.dt(W), R, K, W.stride(0), dead.data_ptr(),
                                                 W32.data_ptr(), K, st), "weights_prepare")

        grouped = groupsize != -1
        ng = 1 if not grouped else (K + groupsize - 1) // groupsize
        gscale = gzero = None
        if static_groups and grouped:                   # gptq.py:157-163, on the un-permuted, un-updated W
            gscale, gzero = find_params_groups(W32, maxq, q.sym, q.mse, q.norm, q.grid, q.maxshrink, groupsize)

        perm = invperm = None
        if actorder:                                    # gptq.py:165-169
            perm = torch.empty(K, device=dev, dtype=torch.int32)
            invperm = torch.empty(K, device=dev, dtype=torch.int32)
            _lib.check(lib.gptq_b200_argsort_desc(diag.data_ptr(), K, perm.data_ptr(), invperm.data_ptr(), st), "argsort")
            Wp = torch.empty_like(W32)
            _lib.check(lib.gptq_b200_permute_cols(W32.data_ptr(), R, K, K, perm.data_ptr(), Wp.data_ptr(), K, st),
                       "permute_cols")
            del W32
        else:
            Wp = W32

        Kp = _round_up(K, 128)                          # gptq.py:174-180
        A = torch.empty((Kp, Kp), **f32)
        _lib.check(lib.gptq_b200_hessian_gather(H.data_ptr(), K, K, _lib.ptr(perm), damp.data_ptr(), A.data_ptr(), Kp, Kp,
                                                st), "hessian_gather")
        del H
        _p0.__exit__()
        # Linears of one subset that see the same activations (q/k/v, gate/up) reach this point with bit-identical A.
        # The chain is a pure function of A, so it is computed once per distinct A: the candidate is compared with the
        # previous input EXACTLY (torch.equal, on device); the memo lives until the next GPTQ object is constructed
        # (= the next calibration round), so nothing is ever reused across layers, subsets or benchmark steps.
        hit = None
        if CHAIN_MEMO:
            ent = _CHAIN_MEMO.get((dev, Kp))
            if ent is not None and torch.equal(A, ent[0]):
                hit = ent
        if hit is not None:
            _, U, info = hit
            del A
        else:
            A_in = A.clone() if CHAIN_MEMO else None
            info = torch.zeros(1, device=dev, dtype=torch.int32)
            nws = lib.gptq_b200_hinv_upper_workspace_bytes(Kp)
            ws = torch.empty(nws, device=dev, dtype=torch.uint8)
            U = torch.empty((Kp, Kp), **f32)
            with _prof("hinv_upper", 2.0 / 3.0 * Kp ** 3):
                _lib.check(lib.gptq_b200_hinv_upper(A.data_ptr(), Kp, Kp, U.data_ptr(), Kp, info.data_ptr(), ws.data_ptr(),
                                                    nws, st), "hinv_upper")
            del ws, A
            if CHAIN_MEMO:
                _CHAIN_MEMO.clear()
                _CHAIN_MEMO[(dev, Kp)] = (A_in, U, info)

        if not q.ready():                               # gptq.py:181-185
            q.find_params(Wp)

        if not grouped:
            mode = 0
            scale = q.scale.to(dev).reshape(R, 1).float().contiguous().clone()
            zero = q.zero.to(dev).reshape(R, 1).float().contiguous().clone()
            gidx = None
        elif static_groups:
            mode = 0
            scale, zero = gscale, gzero
            cols = perm.long() if actorder else torch.arange(K, device=dev)
            gidx = (cols // groupsize).to(torch.int32).contiguous()
        else:
            mode = 1
            scale = torch.zeros((R, ng), **f32)
            zero = torch.zeros((R, ng), **f32)
            gidx = (torch.arange(K, device=dev) // groupsize).to(torch.int32).contiguous()

        Q = torch.empty((R, K), **f32)
        codes = torch.empty((R, K), device=dev, dtype=torch.uint8)
        loss = torch.zeros(1, device=dev, dtype=torch.float64)
        nws = lib.gptq_b200_colloop_workspace_bytes(R, K, blocksize)
        ws = torch.empty(max(nws, 16), device=dev, dtype=torch.uint8)
        # algorithmic bytes of the column loop: W fp16 in, U upper fp32, Q fp16 + codes out
        with _prof("colloop", 2.0 * R * K + 2.0 * K * (K + 1) + 2.0 * R * K + R * K + 8.0 * R * ng):
            _lib.check(lib.gptq_b200_colloop(Wp.data_ptr(), R, K, K, U.data_ptr(), Kp, blocksize, groupsize, mode,
                                             _lib.ptr(gidx), scale.data_ptr(), zero.data_ptr(), ng, maxq, int(q.sym),
                                             int(q.mse), float(q.norm), int(q.grid), float(q.maxshrink), Q.data_ptr(), K,
                                             codes.data_ptr(), K, loss.data_ptr(), ws.data_ptr(), nws, st), "colloop")

        # gptq.py:300-305: un-permute and write back in the layer's dtype
        wdtype = self.layer.weight.data.dtype
        Wout = torch.empty((R, K), device=dev, dtype=wdtype)
        codes_out = torch.empty((R, K), device=dev, dtype=torch.uint8)
        with _prof("epilogue", 4.0 * R * K + R * K + 2.0 * R * K + R * K):
            _lib.check(lib.gptq_b200_finalize(Q.data_ptr(), K, codes.data_ptr(), K, R, K, _lib.ptr(invperm), Wout.data_ptr(),
                                              _lib.dt(Wout), K, codes_out.data_ptr(), K, st), "finalize")

        torch.cuda.synchronize()                        # gptq.py:292
        bad = int(info.item())
        if bad != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not positive-definite "
                f"(the leading minor of order {bad} is not positive-definite).")
        self.error = float(loss.item())
        print('time %.2f' % (time.time() - tick))
        print('error', self.error)

        if row_range is None:
            if isinstance(self.layer, _Conv1D):
                Wfull = Wout.t()
            else:
                Wfull = Wout
            self.layer.weight.data = Wfull.reshape(self.layer.weight.shape).to(wdtype)
        self.Wq = Wout                                  # [rows(shard), K] dequantised weights in the layer dtype
        self.codes = codes_out
        self.scales = scale
        self.zeros = zero
        self.perm = perm
        if gidx is None:
            self.g_idx = torch.zeros(K, device=dev, dtype=torch.int32)
        else:
            self.g_idx = gidx[invperm.long()].contiguous() if actorder else gidx
        # what `self.quantizer` holds after the run: the last group's parameters
        if grouped:
            last = int(self.g_idx[int(perm[K - 1].item())].item()) if (static_groups and actorder) else ng - 1
            if static_groups:
                q = copy.deepcopy(q)                   # the reference swaps in groups[...] (a deepcopy), gptq.py:260
                self.quantizer = q
            q.scale = scale[:, last:last + 1].clone()
            q.zero = zero[:, last:last + 1].clone()
        if DEBUG:
            print(torch.sum((self.layer(self.inp1) - self.out1) ** 2))

    def free(self):
        if DEBUG:
            self.inp1 = None
            self.out1 = None
        self.H = None
        self._stage = None
        self._ws = None
        self._last = None
        self.Losses = None
        self.Trace = None
        # The reference calls torch.cuda.empty_cache() here (gptq.py:318), which forces a cudaFree/cudaMalloc
        # round trip (and a device sync) for every linear; the caching allocator makes it unnecessary.
        if EMPTY_CACHE_ON_FREE:
            torch.cuda.empty_cache()
