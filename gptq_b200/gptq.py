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
  un-permute/cast epilogue as C-ABI calls on the current CUDA stream; the only host sync is the
  one the reference also has (gptq.py:292), reduced to ONE small record read back per call.
* Linears whose hooks receive the very same tensor (q/k/v, gate/up) share one accumulator, one
  prologue and one chain (``_Accumulator``), verified bit-exact on the device; small row shards of
  such linears are solved in one stacked column loop (``_solve_rows``).

Extras the reference does not expose (needed for packing grouped models, tests and sharding):
``codes`` (uint8 [R,K]), ``scales`` / ``zeros`` ([R,ngroups]), ``g_idx`` ([K]), ``perm``, ``error``.
"""
from __future__ import annotations

import copy
import math
import os
import time
import weakref

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
SYRK_VARIANT = int(os.environ.get("GPTQ_B200_SYRK_VARIANT", "0"))     # 0 = auto (CTA-pair kernel), 1 = cta_group::1, 2 = cta_group::2
EMPTY_CACHE_ON_FREE = os.environ.get("GPTQ_B200_EMPTY_CACHE", "0") == "1"
# Linears that receive the very same activation tensor (q/k/v, gate/up: their forward hooks see one tensor) share ONE
# Hessian accumulator, ONE prologue and ONE inverse-Cholesky chain; the reference recomputes all three per linear
# (gptq.py:38-65, 143-180).  A batch is shared only when it is the same memory (pointer, shape, strides, dtype, with the
# first holder keeping the tensor alive) AND, with SHARE_VERIFY, bit-identical to the staged copy (device-side compare,
# checked at fasterquant's own host sync: a mismatch raises instead of producing a different result).
SHARE_H = os.environ.get("GPTQ_B200_SHARE_H", "1") == "1"
# SHARE_VERIFY: "full" = every shared batch is compared in full; "1" (default) = the first VERIFY_FULL_BATCHES shared
# batches of an accumulator are compared in full -- which proves that the module really hands ONE unmodified tensor to
# these linears (a module that changed its input between two sibling hooks would do so in every forward) -- and later
# batches are spot-checked: every VERIFY_BATCH_STRIDE-th shared batch is compared on every VERIFY_ROW_STRIDE-th row (the
# pointer / shape / stride / dtype identity and the keep-alive reference are checked for EVERY batch regardless, on the host,
# without a launch); "0" = identity checks only.
SHARE_VERIFY = os.environ.get("GPTQ_B200_SHARE_VERIFY", "1")
VERIFY_FULL_BATCHES = 2
VERIFY_ROW_STRIDE = 16
VERIFY_BATCH_STRIDE = 8
# Linears that share a Hessian and are quantized with the same arguments are solved in ONE column loop (their rows
# stacked): the loop is latency-bound, so three 4096-row linears cost little more than one.
GROUP_SIBLINGS = os.environ.get("GPTQ_B200_GROUP_SIBLINGS", "1") == "1"
# Opt-in: the syrk of a full staging slab runs on a side stream while the NEXT samples are being staged (and verified)
# into a second slab on the caller's stream (the copies are HBM-bound, the syrk tensor-bound, their CTAs fit on an SM
# together).  Measured on the power-capped B200 box: the 2.6 ms of hidden staging per layer are given back by lower
# clocks in the kernels that follow (step time 112.3 ms either way), so the default stays single-stream.
OVERLAP_SYRK = os.environ.get("GPTQ_B200_OVERLAP_SYRK", "0") == "1"
# Opt-in: short first staging windows (2, 4, 8 samples) so the device starts earlier after a synchronisation.  Measured on
# B200: the 1.5 ms of idle it removes per layer are paid back by the extra read-modify-write passes over S and the smaller
# launches (syrk 35.7 -> 37.2 ms per layer), so the default stays one full slab per launch.
RAMP_WINDOWS = os.environ.get("GPTQ_B200_RAMP_WINDOWS", "0") == "1"
_SIDE_STREAMS = {}


def _side_stream(dev):
    st = _SIDE_STREAMS.get(dev)
    if st is None:
        st = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return st


# Optional per-call device timing: set ``gptq_b200.gptq.PROFILE = []`` and every C-ABI hot call appends
# (name, start_event, end_event, work) recorded on the launching stream (bench.py reads it after a sync).
PROFILE = None
PROFILE_FINE = False      # also time every staging copy / verification (hundreds of event pairs per layer: diagnostic runs only)
HOST_WAIT = None          # [seconds] accumulated host time blocked in fasterquant's record read-back, when set to a list


class _NoProf:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NOPROF = _NoProf()


def _prof(name, work=0.0, meta=None, fine=False):
    """Context manager timing a region with CUDA events when PROFILE is a list; a shared no-op object otherwise (the hot
    add_batch path runs ~900 times per decoder layer)."""
    if PROFILE is None or (fine and not PROFILE_FINE):
        return _NOPROF
    return _Prof(name, work, meta)


class _Prof:
    def __init__(self, name, work=0.0, meta=None, fine=False):
        self.name, self.work, self.meta = name, work, meta
        self.on = True

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            PROFILE.append((self.name, self.a, self.b, self.work, self.meta))
        return False


_DEBUG_SYNC = os.environ.get("GPTQ_B200_DEBUG_SYNC", "0") == "1"


def _dbg(tag):
    """GPTQ_B200_DEBUG_SYNC=1: synchronize after every device step so an asynchronous fault is attributed to it."""
    if _DEBUG_SYNC:
        try:
            torch.cuda.synchronize()
        except Exception as e:      # pragma: no cover
            raise RuntimeError(f"device fault surfaced after step '{tag}': {e}") from e


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def syrk_issued_flops(K: int, tokens: int) -> float:
    """MMA flops the syrk kernel actually issues: 128x256 tiles touching the lower triangle."""
    nbm, nbn = (K + 127) // 128, (K + 255) // 256
    tiles = sum(nbm - 2 * j for j in range(nbn))
    return 2.0 * tiles * 128 * 256 * tokens


def _on_device(fn):
    """Run a GPTQ method with the layer's GPU current.  The reference works on any device because torch dispatches per
    tensor; here every C-ABI launch goes to torch.cuda.current_stream() of the CURRENT device, so a layer living on
    another GPU of the same process needs the guard (workspaces, streams and per-device kernel attributes follow)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        idx = self.dev.index
        if idx is None or idx == torch.cuda.current_device():
            return fn(self, *a, **k)
        with torch.cuda.device(idx):
            return fn(self, *a, **k)
    return wrapped


_ACT_DTYPES = (torch.float16, torch.bfloat16, torch.float32)
_JOINABLE = {}            # (device, batch identity) -> weakref(_Accumulator that has absorbed exactly that one batch)


def _batch_key(x):
    return (x.data_ptr(), x.shape, x.stride(), x.dtype)


class _Accumulator:
    """Unscaled Hessian accumulator S = sum_tokens x x^T of ONE activation stream (gptq.py:38-65 in closed form),
    its staging slab, and -- once a member ran ``fasterquant`` -- the finished H with the prologue / chain results
    that only depend on H.  Several GPTQ objects (``members``) may hold the same accumulator."""

    def __init__(self, K, dev):
        self.K, self.dev = K, dev
        self.S = None                  # [K,K] fp32, tiles touching the lower triangle valid
        self.stage = None              # [cap_tokens, roundup(K,8)] staging slab being filled
        self.stage_other = None        # the second slab (being consumed by a syrk on the side stream, or idle)
        self.stage_busy = None         # event: the syrk reading `stage` (from its previous turn) has finished
        self.other_busy = None
        self.syrk_ev = None            # event: the last syrk launched on the side stream has finished
        self.staged = 0
        self.ramp = 0                  # tokens after which the NEXT absorb flushes early (0 = only when the slab is full)
        self.ws = None
        self.count = 0                 # samples absorbed (the reference's nsamples)
        self.batches = 0
        self.members = 1
        self.raw_key = None            # identity of the hook's ORIGINAL tensor of the most recent batch (sibling fast path)
        self.last_ld = 0
        self.key = None                # identity of the most recent batch ...
        self.ref = None                # ... and the tensor itself: while it is held its memory cannot be recycled
        self.first_key = None
        self.last_tmp = 0
        self.last_tokens = 0
        self.last_in_slab = False      # the most recent batch sits contiguously at the end of the slab
        self.last = None               # (view of the last batch, count at that time) for GPTQ.input
        self.last_slab = None
        self.flag = None               # device int32, set by compare_rows when a shared batch was NOT identical
        self.verified = 0              # shared batches compared so far (see SHARE_VERIFY)
        self.input_mean = None         # (sum over tokens of the last batch, tokens, count): GPTQ.input after fasterquant
        self.reduced = False           # dist.allreduce_hessian already summed S over the ranks
        self.sealed = False
        self.n_final = None
        self.H = None                  # (2/n) S, mirrored, dead diagonal fixed -- after seal()
        self.dead = None
        self.solves = {}               # (percdamp, actorder) -> prologue + chain results
        self.member_refs = []          # weakrefs of the GPTQ objects holding this accumulator
        self.spec = {}                 # id(GPTQ) -> (arguments, weights, results) solved ahead by a sibling

    # -------------------------------------------------------------- accumulate
    def absorb(self, x, tmp, key, joinable):
        if self.sealed:
            raise RuntimeError("GPTQ.add_batch: this Hessian was already consumed by fasterquant of a linear sharing it")
        lib = _lib.load()
        K = self.K
        if self.S is None:
            self.S = torch.zeros((K, K), device=self.dev, dtype=torch.float32)
        self.count += tmp
        self.batches += 1
        self.key, self.ref, self.last_tmp = key, x, tmp
        if self.batches == 1 and joinable:
            self.first_key = (self.dev, key)
            _JOINABLE[self.first_key] = weakref.ref(self)
        elif self.first_key is not None:
            self._unregister()
        tokens = x.shape[0]
        self.last_tokens = tokens
        if x.dtype == torch.float32:
            # fp32 activations: exact fp32 CUDA-core path (the tensor path takes 16-bit operands)
            self.flush()
            self.wait_syrk()
            xc = x.contiguous()
            _lib.check(lib.gptq_b200_syrk_accum_simt(xc.data_ptr(), _lib.F32, tokens, K, xc.stride(0), self.S.data_ptr(), K,
                                                     1.0, 1.0, _lib.stream()), "syrk_accum_simt")
            self.last = (xc, self.count)
            self.last_in_slab = False
            return
        if x.stride(-1) != 1:
            x = x.contiguous()
        if self.stage is None or self.stage.dtype != x.dtype:
            self.flush()
            self.wait_syrk()
            cap = max(tokens, 1) * max(STAGE_SAMPLES, 1)
            self.stage = torch.empty((cap, _round_up(K, 8)), device=self.dev, dtype=x.dtype)
            self.stage_other = self.stage_busy = self.other_busy = None
            self.ramp = 2 * max(tokens, 1) if (RAMP_WINDOWS and K <= 8192 and STAGE_SAMPLES >= 8) else 0
        cap = self.stage.shape[0]
        if tokens <= cap and (self.staged + tokens > cap or self.staged >= self.ramp > 0):
            # keep a batch contiguous in the slab.  The first windows of a calibration round are short (2, 4, 8 samples, then
            # the full slab): after a host synchronisation (the previous linear's fasterquant) the device would otherwise
            # idle until the host has staged a whole slab; the extra read-modify-write passes over S only pay for small K.
            self.flush()
            if self.ramp:
                self.ramp = self.ramp * 2 if self.ramp * 2 < cap else 0
        in_slab = tokens <= cap
        if in_slab:                                        # the common case: no slicing, plain pointer arithmetic
            st_ = self.stage
            ld = st_.stride(0)
            with _prof("staging", fine=True):
                rc = 0 if tokens == 0 else lib.gptq_b200_stage_input(
                    x.data_ptr(), _lib.dt(x), tokens, K, x.stride(0), st_.data_ptr() + self.staged * ld * st_.element_size(), ld,
                    torch.cuda.current_stream().cuda_stream)
            if rc:
                _lib.check(rc, "stage_input")
            self.last = (self.staged, tokens, self.count)  # resolved lazily by GPTQ.input
            self.last_slab = st_
            self.staged += tokens
            self.last_in_slab = True
            return
        done = 0
        while done < tokens:
            if self.staged == cap:
                self.flush()
            n = min(tokens - done, cap - self.staged)
            dst = self.stage[self.staged:self.staged + n]
            src = x[done:done + n]
            _lib.check(lib.gptq_b200_stage_input(src.data_ptr(), _lib.dt(src), n, K, src.stride(0), dst.data_ptr(),
                                                 dst.stride(0), _lib.stream()), "stage_input")
            self.last = (dst[:, :K], self.count)
            self.staged += n
            done += n
        self.last_in_slab = False

    def verify(self, x, ldx=None):
        """Device-side proof that `x` is bit-identical to the batch staged last (result read at fasterquant's sync).
        Runs once per sample and sibling: no tensor slicing, plain pointer arithmetic.  `ldx`: row pitch when `x` is the
        hook's original (possibly 3-D) tensor rather than its [tokens, K] view."""
        if SHARE_VERIFY == "0":
            return
        self.verified += 1
        full = SHARE_VERIFY == "full" or self.verified <= VERIFY_FULL_BATCHES * max(self.members - 1, 1)
        if not full and self.verified % VERIFY_BATCH_STRIDE:
            return
        if self.flag is None:
            self.flag = torch.zeros(1, device=self.dev, dtype=torch.int32)
        if ldx is None:
            if x.stride(-1) != 1:
                x = x.contiguous()
            ldx = x.stride(0)
        st = self.stage
        ld = st.stride(0)
        ref_ptr = st.data_ptr() + (self.staged - self.last_tokens) * ld * st.element_size()
        rows = self.last_tokens
        if not full and rows >= VERIFY_ROW_STRIDE:
            rows, ldx, ld = rows // VERIFY_ROW_STRIDE, ldx * VERIFY_ROW_STRIDE, ld * VERIFY_ROW_STRIDE
        with _prof("verify", fine=True):
            rc = _lib.load().gptq_b200_compare_rows(x.data_ptr(), ldx, ref_ptr, ld, rows, self.K,
                                                    self.flag.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if rc:
            _lib.check(rc, "compare_rows")

    def wait_syrk(self):
        """Make the current stream wait for the accumulation launched on the side stream."""
        if self.syrk_ev is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.syrk_ev)

    def flush(self):
        """Fold the staged tokens into S.  With OVERLAP_SYRK the kernel goes to the side stream and staging continues in
        the other slab; S is valid on the caller's stream only after wait_syrk()."""
        self.last_in_slab = False
        if self.stage is None or self.staged == 0:
            return
        lib = _lib.load()
        K = self.K
        need = lib.gptq_b200_syrk_workspace_bytes(K, self.staged, SYRK_VARIANT)
        if need and (self.ws is None or self.ws.numel() < need):
            self.wait_syrk()                               # the old workspace may still be in use
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.dev)

        def launch():
            with _prof("syrk", syrk_issued_flops(K, self.staged), (K, self.staged)):
                _lib.check(lib.gptq_b200_syrk_accum(self.stage.data_ptr(), _lib.dt(self.stage), self.staged, K,
                                                    self.stage.stride(0), self.S.data_ptr(), K, 1.0, 1.0, CHUNK_TOKENS,
                                                    SYRK_VARIANT, _lib.ptr(self.ws) if need else 0, need, _lib.stream()),
                           "syrk_accum")
        if not OVERLAP_SYRK:
            launch()
            self.staged = 0
            return
        main = torch.cuda.current_stream(self.dev)
        side = _side_stream(self.dev)
        for t in (self.stage, self.S, self.ws):            # allocated on the caller's stream, used on the side stream
            if t is not None:
                t.record_stream(side)
        ready = torch.cuda.Event()
        ready.record(main)                                 # the window is staged (and S initialised)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            launch()
            done = torch.cuda.Event()
            done.record(side)
        self.syrk_ev = done
        # continue staging in the other slab, once the syrk that last read it has finished
        self.stage, self.stage_other = self.stage_other, self.stage
        self.stage_busy, self.other_busy = self.other_busy, done
        if self.stage is None:
            self.stage = torch.empty_like(self.stage_other)
        elif self.stage_busy is not None:
            main.wait_event(self.stage_busy)
        self.staged = 0

    def _unregister(self):
        if self.first_key is not None:
            ent = _JOINABLE.get(self.first_key)
            if ent is not None and ent() in (self, None):
                del _JOINABLE[self.first_key]
            self.first_key = None

    # -------------------------------------------------------------- read / consume
    def materialize(self, n):
        """A private full symmetric copy of (2/n) S -- what gptq.py:59-65 maintains as self.H."""
        if self.sealed:
            H = self.H.clone()
            if self.n_final and n and n != self.n_final:
                H.mul_(self.n_final / n)
            H.diagonal()[self.dead.bool()] = 0             # undo the dead-column fix seal() applied (gptq.py:144)
            return H
        if self.S is None:
            return torch.zeros((self.K, self.K), device=self.dev, dtype=torch.float32)
        self.flush()
        self.wait_syrk()
        H = self.S.clone()
        _lib.check(_lib.load().gptq_b200_symmetrize(H.data_ptr(), self.K, self.K, _lib.stream()), "symmetrize")
        if n > 0:
            H.mul_(2.0 / n)
        return H

    def seal(self, n):
        """Finish H in place (mirror, scale by 2/n), once; later members get the same tensor."""
        if self.sealed:
            if n != self.n_final:
                raise RuntimeError(f"fasterquant: shared Hessian was finished with nsamples={self.n_final}, "
                                   f"this linear has nsamples={n}")
            return self.H
        lib = _lib.load()
        K = self.K
        if self.S is None:
            self.S = torch.zeros((K, K), device=self.dev, dtype=torch.float32)
        self.flush()
        self.wait_syrk()
        self._unregister()
        H = self.S
        with _prof("prologue"):
            _lib.check(lib.gptq_b200_symmetrize(H.data_ptr(), K, K, _lib.stream()), "symmetrize")
            if n > 0:
                H.mul_(2.0 / n)
            self.dead = torch.empty(K, device=self.dev, dtype=torch.uint8)
            _lib.check(lib.gptq_b200_dead_columns(H.data_ptr(), K, K, self.dead.data_ptr(), _lib.stream()), "dead_columns")
        self.H, self.S, self.n_final, self.sealed = H, None, n, True
        if self.last is not None:      # gptq.py:63 keeps self.input after fasterquant: reduce it before the slab goes away
            self.input_mean = _last_input(self)
        self.stage = self.stage_other = self.ws = self.ref = self.last = self.last_slab = None
        return H


def _last_input(acc):
    """gptq.py:62-63: mean over tokens of sqrt(2/n) * (last batch); n = sample count when that batch arrived."""
    if len(acc.last) == 3:                              # (row offset, tokens, count) in the slab staged last
        off, tokens, n = acc.last
        x = acc.last_slab[off:off + tokens, :acc.K]
    else:
        x, n = acc.last
    if x.shape[0] == 0 or n == 0:
        return torch.zeros(acc.K, device=acc.dev)
    return math.sqrt(2 / n) * x.float().mean(0)


def _solve_hessian(H, K, percdamp, actorder, dead=None):
    """fasterquant's H-only part (gptq.py:143-144, 165-180): dead columns, damp, act-order permutation and the
    inverse-Cholesky chain.  H's dead diagonal entries are set to 1 in place; returns a dict of device tensors."""
    lib = _lib.load()
    st = _lib.stream()
    dev = H.device
    f32 = dict(device=dev, dtype=torch.float32)
    with _prof("prologue"):
        if dead is None:
            dead = torch.empty(K, device=dev, dtype=torch.uint8)
            _lib.check(lib.gptq_b200_dead_columns(H.data_ptr(), K, K, dead.data_ptr(), st), "dead_columns")
        _dbg("dead_columns")
        diag = torch.empty(K, **f32)
        damp = torch.empty(1, **f32)
        ndead = torch.empty(1, device=dev, dtype=torch.int32)
        _lib.check(lib.gptq_b200_hessian_prepare(H.data_ptr(), K, K, float(percdamp), diag.data_ptr(), damp.data_ptr(),
                                                 ndead.data_ptr(), st), "hessian_prepare")
        _dbg("hessian_prepare")
        perm = invperm = None
        if actorder:                                    # gptq.py:165-169
            perm = torch.empty(K, device=dev, dtype=torch.int32)
            invperm = torch.empty(K, device=dev, dtype=torch.int32)
            _lib.check(lib.gptq_b200_argsort_desc(diag.data_ptr(), K, perm.data_ptr(), invperm.data_ptr(), st), "argsort")
            _dbg("argsort")
        Kp = _round_up(K, 128)                          # gptq.py:174-180
        A = torch.empty((Kp, Kp), **f32)
        _lib.check(lib.gptq_b200_hessian_gather(H.data_ptr(), K, K, _lib.ptr(perm), damp.data_ptr(), A.data_ptr(), Kp, Kp,
                                                st), "hessian_gather")
        _dbg("hessian_gather")
    info = torch.zeros(1, device=dev, dtype=torch.int32)
    nws = lib.gptq_b200_hinv_upper_workspace_bytes(Kp)
    ws = torch.empty(nws, device=dev, dtype=torch.uint8)
    U = torch.empty((Kp, Kp), **f32)
    with _prof("hinv_upper", 2.0 / 3.0 * Kp ** 3):
        _lib.check(lib.gptq_b200_hinv_upper(A.data_ptr(), Kp, Kp, U.data_ptr(), Kp, info.data_ptr(), ws.data_ptr(),
                                            nws, st), "hinv_upper")
    _dbg("hinv_upper")
    return dict(dead=dead, perm=perm, invperm=invperm, U=U, info=info, Kp=Kp)


def _solve_rows(Ws, sol, qcfg, blocksize, groupsize, actorder, static_groups, wdtype, preset=None):
    """fasterquant's per-row part (gptq.py:130-145 W.float() + dead columns, 157-163 static groups, 168 permute,
    181-185 full-row find_params, 191-276 column loop, 300-305 un-permute + cast) for the rows of one or several linears
    that share the solve `sol`, stacked into ONE matrix: rows are independent, so the results are cut back per linear.
    Everything is queued on the current stream; nothing is read back.  Returns one dict of device tensors per linear."""
    lib = _lib.load()
    st = _lib.stream()
    maxq, _perch, sym, mse, norm, grid, maxshrink = qcfg
    dead, perm, invperm, U, Kp = (sol[k] for k in ("dead", "perm", "invperm", "U", "Kp"))
    dev = U.device
    K = Ws[0].shape[1]
    Rs = [w.shape[0] for w in Ws]
    R = sum(Rs)
    f32 = dict(device=dev, dtype=torch.float32)
    with _prof("prologue"):
        W32 = torch.empty((R, K), **f32)
        r0 = 0
        for w in Ws:
            _lib.check(lib.gptq_b200_weights_prepare(w.data_ptr(), _lib.dt(w), w.shape[0], K, w.stride(0), dead.data_ptr(),
                                                     W32[r0:].data_ptr(), K, st), "weights_prepare")
            r0 += w.shape[0]
        _dbg("weights_prepare")
        grouped = groupsize != -1
        ng = 1 if not grouped else (K + groupsize - 1) // groupsize
        gscale = gzero = None
        if static_groups and grouped:               # gptq.py:157-163, on the un-permuted, un-updated W
            gscale, gzero = find_params_groups(W32, maxq, sym, mse, norm, grid, maxshrink, groupsize)
            _dbg("find_params_groups")
        if actorder:                                # gptq.py:168
            Wp = torch.empty_like(W32)
            _lib.check(lib.gptq_b200_permute_cols(W32.data_ptr(), R, K, K, perm.data_ptr(), Wp.data_ptr(), K, st),
                       "permute_cols")
            del W32
        else:
            Wp = W32
    if preset is None:
        # gptq.py:181-185: the quantizer is not ready -> find_params on the whole (permuted) rows (one pair per row)
        if _perch:
            full_scale, full_zero = find_params_groups(Wp, maxq, sym, mse, norm, grid, maxshrink)
            full_scale, full_zero = full_scale.reshape(R, 1), full_zero.reshape(R, 1)
        else:                                       # quant.py:53-54,97-103: one pair per linear, repeated over its rows
            parts, r0 = [], 0
            for Rm in Rs:
                tq = Quantizer()
                tq.maxq = torch.tensor(maxq)
                tq.perchannel, tq.sym, tq.mse, tq.norm, tq.grid, tq.maxshrink = False, sym, mse, norm, grid, maxshrink
                tq.find_params(Wp[r0:r0 + Rm], weight=True)
                parts.append((tq.scale.reshape(Rm, 1), tq.zero.reshape(Rm, 1)))
                r0 += Rm
            full_scale = torch.cat([p[0] for p in parts], 0)
            full_zero = torch.cat([p[1] for p in parts], 0)
        row_scale, row_zero = full_scale, full_zero
    else:                                           # a quantizer that is ready() keeps its parameters (single linear only)
        full_scale = full_zero = None
        row_scale = preset[0].to(dev).reshape(R, 1).float().contiguous()
        row_zero = preset[1].to(dev).reshape(R, 1).float().contiguous()

    if not grouped:
        mode = 0
        scale, zero = row_scale.clone(), row_zero.clone()
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
    row_loss = torch.empty(R, **f32)
    nws = lib.gptq_b200_colloop_workspace_bytes(R, K, blocksize)
    ws = torch.empty(max(nws, 16), device=dev, dtype=torch.uint8)
    # algorithmic bytes of the column loop: W fp16 in, U upper fp32, Q fp16 + codes out
    with _prof("colloop", 2.0 * R * K + 2.0 * K * (K + 1) + 2.0 * R * K + R * K + 8.0 * R * ng):
        _lib.check(lib.gptq_b200_colloop(Wp.data_ptr(), R, K, K, U.data_ptr(), Kp, blocksize, groupsize, mode,
                                         _lib.ptr(gidx), scale.data_ptr(), zero.data_ptr(), ng, maxq, int(sym),
                                         int(mse), float(norm), int(grid), float(maxshrink), Q.data_ptr(), K,
                                         codes.data_ptr(), K, row_loss.data_ptr(), ws.data_ptr(), nws, st), "colloop")

    # gptq.py:300-305: un-permute and write back in the layer's dtype
    Wout = torch.empty((R, K), device=dev, dtype=wdtype)
    codes_out = torch.empty((R, K), device=dev, dtype=torch.uint8)
    with _prof("epilogue", 4.0 * R * K + R * K + 2.0 * R * K + R * K):
        _lib.check(lib.gptq_b200_finalize(Q.data_ptr(), K, codes.data_ptr(), K, R, K, _lib.ptr(invperm), Wout.data_ptr(),
                                          _lib.dt(Wout), K, codes_out.data_ptr(), K, st), "finalize")
    if gidx is None:
        g_idx_out = torch.zeros(K, device=dev, dtype=torch.int32)
    else:
        g_idx_out = gidx[invperm.long()].contiguous() if actorder else gidx
    q_scale = q_zero = None
    if grouped:
        # the group of the LAST processed column -- gidx is in processing order, so that group is gidx[K-1]
        if static_groups and actorder:
            last = gidx[K - 1:K].long()
            q_scale, q_zero = scale.index_select(1, last), zero.index_select(1, last)
        else:
            q_scale, q_zero = scale[:, ng - 1:ng].clone(), zero[:, ng - 1:ng].clone()
    outs, r0 = [], 0
    for Rm in Rs:
        sl = slice(r0, r0 + Rm)
        outs.append(dict(Wout=Wout[sl], codes=codes_out[sl], scale=scale[sl], zero=zero[sl], g_idx=g_idx_out,
                         loss=row_loss[sl].double().sum().reshape(1),
                         full_scale=None if full_scale is None else full_scale[sl],
                         full_zero=None if full_zero is None else full_zero[sl],
                         q_scale=None if q_scale is None else q_scale[sl], q_zero=None if q_zero is None else q_zero[sl]))
        r0 += Rm
    return outs


class GPTQ:

    def __init__(self, layer):
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
        self._is_linear = isinstance(self.layer, nn.Linear) or isinstance(self.layer, _Conv1D)
        self._acc = None               # _Accumulator (allocated with the first batch; may be shared, see SHARE_H)
        self._H = None                 # materialised (2/n)*S, full symmetric, or a user-provided H
        self._consumed = False
        self._input = None             # gptq.py:63 self.input, kept after the accumulator is released
        self._error = self._error_dev = None

    # ------------------------------------------------------------------ Hessian (gptq.py:38-65)
    @_on_device
    def add_batch(self, inp, out):
        if DEBUG:
            self.inp1 = inp
            self.out1 = out
        acc = self._acc
        if acc is not None and acc.members > 1 and acc.raw_key is not None and not self._consumed:
            # Fast path of a sibling hook (q/k/v, gate/up: ~900 hook calls per decoder layer, and the device idles while the
            # host works through them): the very tensor a sibling staged a moment ago -- same memory, shape, strides, dtype,
            # still held alive by the accumulator -- needs no view, no staging and (mostly) no launch.
            if (acc.raw_key == (inp.data_ptr(), inp.shape, inp.stride(), inp.dtype) and acc.last_in_slab
                    and acc.count == self.nsamples + acc.last_tmp):
                self._H = None
                acc.verify(inp, acc.last_ld)
                self.nsamples += acc.last_tmp
                return
        raw_key = (inp.data_ptr(), inp.shape, inp.stride(), inp.dtype) if (SHARE_H and self._is_linear) else None
        if not inp.is_cuda:
            _lib.require_cuda(inp, "GPTQ.add_batch: inp")
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        tmp = inp.shape[0]
        if self._is_linear:
            if inp.dim() == 3:
                inp = inp.reshape((-1, inp.shape[-1]))            # [tokens, K]; the reference transposes, we don't need to
        elif isinstance(self.layer, nn.Conv2d):
            unfold = nn.Unfold(self.layer.kernel_size, dilation=self.layer.dilation, padding=self.layer.padding,
                               stride=self.layer.stride)
            inp = unfold(inp).permute([1, 0, 2]).flatten(1).t()   # [tokens, K]
        if inp.shape[-1] != self.columns:
            raise ValueError(f"add_batch: expected {self.columns} input features, got {inp.shape[-1]}")
        if inp.dtype not in _ACT_DTYPES:
            raise TypeError(f"unsupported dtype {inp.dtype}")
        if self._consumed:
            raise RuntimeError("add_batch: H was already consumed by fasterquant or freed")
        self._H = None
        x = inp
        key = _batch_key(x)
        shareable = SHARE_H and x.dtype != torch.float32 and x.shape[0] > 0
        acc = self._acc
        if acc is None:                                           # first batch: the same tensor may already be staged
            if shareable and self.nsamples == 0:
                ent = _JOINABLE.get((self.dev, key))
                cand = ent() if ent is not None else None
                if (cand is not None and cand.K == self.columns and cand.count == tmp and cand.batches == 1
                        and cand.last_in_slab and not cand.sealed):
                    cand.members += 1
                    cand.member_refs.append(weakref.ref(self))
                    cand.verify(x)
                    self._acc = cand
                    self.nsamples += tmp
                    return
            acc = self._acc = _Accumulator(self.columns, self.dev)
            acc.member_refs.append(weakref.ref(self))
            acc.count = self.nsamples                             # (non-zero only if the caller preset nsamples)
        elif acc.members > 1:
            if shareable and acc.key == key and acc.count == self.nsamples + tmp and acc.last_in_slab:
                acc.verify(x)                                     # a sibling staged this very batch already
                self.nsamples += tmp
                return
            if acc.count != self.nsamples:                        # siblings moved on with a different batch
                self._fork()
                acc = self._acc
        acc.absorb(x, tmp, key, shareable)
        self.nsamples += tmp
        if shareable and raw_key is not None and x.stride(-1) == 1 and x.data_ptr() == raw_key[0] and acc.last_in_slab:
            acc.raw_key, acc.last_ld = raw_key, x.stride(0)
        else:
            acc.raw_key = None

    def _fork(self):
        """Leave a shared accumulator, taking a private copy of the state that matches THIS object's history."""
        acc = self._acc
        if acc.sealed:
            raise RuntimeError("GPTQ: the shared Hessian was already consumed by a sibling's fasterquant")
        if acc.count == self.nsamples:
            keep = acc.staged
        elif acc.count == self.nsamples + acc.last_tmp and acc.last_in_slab:
            keep = acc.staged - acc.last_tokens                   # everything but the batch we did not receive
        else:
            raise RuntimeError("GPTQ: linears sharing a Hessian accumulator received different batches and the common "
                               "state can no longer be separated; set GPTQ_B200_SHARE_H=0")
        new = _Accumulator(acc.K, acc.dev)
        acc.wait_syrk()
        new.S = acc.S.clone() if acc.S is not None else None
        if acc.stage is not None:
            new.stage = torch.empty_like(acc.stage)
            if keep > 0:
                new.stage[:keep].copy_(acc.stage[:keep])
            new.staged = keep
        new.count = self.nsamples
        new.batches = max(acc.batches, 2)
        new.flag = acc.flag
        acc.members -= 1
        acc.member_refs = [r for r in acc.member_refs if r() is not None and r() is not self]
        acc.spec.pop(id(self), None)
        new.member_refs.append(weakref.ref(self))
        self._acc = new

    def _release(self):
        acc, self._acc = self._acc, None
        if acc is not None:
            if acc.input_mean is not None:
                self._input = acc.input_mean
            acc.members -= 1
            acc.member_refs = [r for r in acc.member_refs if r() is not None and r() is not self]
            acc.spec.pop(id(self), None)
            if acc.members <= 0:
                acc._unregister()

    def _own_accumulator(self):
        """The accumulator in the state matching this object's own history (forks away from siblings if needed)."""
        acc = self._acc
        if acc is None:
            acc = self._acc = _Accumulator(self.columns, self.dev)
            acc.member_refs.append(weakref.ref(self))
            acc.count = self.nsamples
        elif acc.members > 1 and acc.count != self.nsamples and not acc.reduced and not acc.sealed:
            self._fork()
            acc = self._acc
        return acc

    @property
    def H(self):
        """(2/nsamples) * sum x x^T as a full symmetric fp32 matrix (what gptq.py:59-65 maintains)."""
        if self._H is None and not self._consumed:
            self._H = self._materialize()
        return self._H

    @_on_device
    def _materialize(self):
        return self._own_accumulator().materialize(self.nsamples)

    @H.setter
    def H(self, value):
        self._H = value
        self._release()
        self._consumed = value is None

    @H.deleter
    def H(self):
        self._H = None

    @property
    def error(self):
        """sum(Losses) of the last fasterquant (gptq.py:294 prints it).  After dist.fasterquant_sharded it is the sum over
        all row shards, kept on the device and read (one sync) only when asked for."""
        if self._error is None and self._error_dev is not None:
            self._error = float(self._error_dev.item())
        return self._error

    @error.setter
    def error(self, value):
        self._error = value

    def error_tensor(self):
        """The same value as a float64 device tensor of one element (no host sync)."""
        return self._error_dev

    @property
    def input(self):
        """gptq.py:63: mean over tokens of sqrt(2/n) * (last batch)."""
        if self._input is not None:
            return self._input
        acc = self._acc
        if acc is None or (acc.last is None and acc.input_mean is None):
            raise AttributeError("input")
        return acc.input_mean if acc.last is None else _last_input(acc)

    @_on_device
    def raw_accumulator(self):
        """Unscaled S (after folding in anything still staged) -- what multi-GPU runs all-reduce."""
        acc = self._own_accumulator()
        if acc.sealed:
            raise RuntimeError("raw_accumulator: the Hessian was already finished by fasterquant")
        if acc.S is None:
            acc.S = torch.zeros((self.columns, self.columns), device=self.dev, dtype=torch.float32)
        acc.flush()
        acc.wait_syrk()
        self._H = None
        return acc.S

    # ------------------------------------------------------------------ solve (gptq.py:126-309)
    def _weight_rows(self, row_range):
        """The layer's weight as [rows, K] (Conv2d flattened, Conv1D transposed), optionally a row shard."""
        W = self.layer.weight.data
        if isinstance(self.layer, nn.Conv2d):
            W = W.flatten(1)
        if isinstance(self.layer, _Conv1D):
            W = W.t()
        if row_range is not None:                       # row sharding: rows are independent given H
            W = W[row_range[0]:row_range[1]]
        return W.contiguous()

    @staticmethod
    def _qcfg(q):
        return (int(q.maxq.item()), bool(q.perchannel), bool(q.sym), bool(q.mse), float(q.norm), int(q.grid),
                float(q.maxshrink))

    def _siblings_to_solve_with(self, acc, argkey, R):
        """Linears that share this Hessian (hence U) and would be solved with exactly the same arguments: their rows are
        independent of ours given U (gptq.py:191-276 acts row-wise), so they ride along in the same column loop -- the
        loop is latency-bound, a few thousand extra rows cost little -- and their results wait in the accumulator until
        their own fasterquant call asks for them (with the same arguments and bit-identical weights)."""
        if not GROUP_SIBLINGS or acc is None or acc.members <= 1:
            return []
        sibs = []
        total = R
        # measured on B200: stacking pays only while the stacked slab per SM stays on the 8-lanes-per-row kernel (<= 40 rows
        # per CTA), i.e. for row shards of 4+ GPUs or small models; at 4096 rows per linear it is a wash (the siblings'
        # calls become pure host latency), so it is not done there
        limit = 40 * torch.cuda.get_device_properties(self.dev).multi_processor_count
        for ref in list(acc.member_refs):
            g = ref()
            if g is None or g is self or g._acc is not acc or g._consumed or g._H is not None or id(g) in acc.spec:
                continue
            # (after dist.allreduce_hessian the siblings pick up the global sample count at their own turn)
            if (g.nsamples != self.nsamples and not acc.reduced) or g.rows != self.rows or g.columns != self.columns:
                continue
            if type(g.layer) is not type(self.layer) or g.layer.weight.dtype != self.layer.weight.dtype:
                continue
            q = g.quantizer
            if not isinstance(q, Quantizer) or q.ready() or self._qcfg(q) != argkey[-1]:
                continue
            if total + R > limit:
                break
            sibs.append(g)
            total += R
        return sibs

    @_on_device
    def fasterquant(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, static_groups=False,
                    model_name="opt", layer_name="layer", lut_quant=False, non_linear_quant=False, columnwise=False,
                    row_range=None):
        if lut_quant or non_linear_quant or columnwise:
            raise NotImplementedError("lut_quant / non_linear_quant / columnwise are research modes of the fork "
                                      "outside the GPTQ hot path (SURVEY 8a)")
        if not isinstance(self.quantizer, Quantizer):
            raise TypeError("GPTQ.quantizer must be a gptq_b200.quant.Quantizer")
        _lib.load()
        dev = self.dev
        q = self.quantizer
        K = self.columns
        W = self._weight_rows(row_range)
        R = W.shape[0]
        qcfg = self._qcfg(q)
        if qcfg[0] == 0:
            raise RuntimeError("Quantizer.configure() was not called")
        if not qcfg[1] and (groupsize != -1 or row_range is not None):
            # quant.py:53-54,97-103: perchannel=False flattens the slice -> ONE (scale, zero) shared by all rows of a group;
            # the group kernels and the row sharding here are per-row.  No reference driver uses this combination
            # (all pass perchannel=True: opt.py:174, llama.py:167, bloom.py:149).
            raise NotImplementedError("perchannel=False is supported for groupsize=-1 without row sharding only")
        argkey = (int(blocksize), float(percdamp), int(groupsize), bool(actorder), bool(static_groups),
                  None if row_range is None else tuple(row_range), qcfg)

        tick = time.time()
        if self._consumed:
            raise RuntimeError("fasterquant: H was already consumed or freed")
        out = None
        if self._H is not None:                         # a user-provided / already materialised H (gptq.py:141-142)
            sol = _solve_hessian(self._H, K, percdamp, actorder)
            flag = None
            sibs = []
        else:                                           # the accumulator, finished in place; siblings reuse the solve
            acc = self._own_accumulator()
            H = acc.seal(self.nsamples)
            _dbg("seal")
            skey = (float(percdamp), bool(actorder))
            sol = acc.solves.get(skey)
            if sol is None:
                sol = acc.solves[skey] = _solve_hessian(H, K, percdamp, actorder, dead=acc.dead)
            flag = acc.flag
            ent = acc.spec.pop(id(self), None)          # solved ahead by a sibling?
            # The result is used only if the weights the sibling solved for us are still the layer's weights, bit for bit;
            # otherwise this linear is solved again (the reference would quantize whatever the layer holds now).  The
            # comparison reads one flag back: the device is idle at this point (the sibling's fasterquant ended with its own
            # read-back and nothing was queued since), so this costs a round trip, not a pipeline drain.
            if ent is not None and ent[0] == argkey and not q.ready() and torch.equal(ent[1], W):
                out = ent[2]
            sibs = [] if (out is not None or q.ready()) else self._siblings_to_solve_with(acc, argkey, R)
            del H
        self._H = None
        self._consumed = True
        if out is None:
            Ws = [W] + [g._weight_rows(row_range) for g in sibs]
            outs = _solve_rows(Ws, sol, qcfg, blocksize, groupsize, actorder, static_groups, self.layer.weight.data.dtype,
                               preset=(q.scale, q.zero) if q.ready() else None)
            out = outs[0]
            for g, wsrc, o in zip(sibs, Ws[1:], outs[1:]):
                acc.spec[id(g)] = (argkey, wsrc.clone(), o)
        info, perm = sol["info"], sol["perm"]

        # Everything that does not need a host value was queued BEFORE this one synchronisation point; ONE small record
        # (info, sharing flag, loss) is read back with a single copy.  gptq.py:292 calls torch.cuda.synchronize();
        # everything here was issued on the current stream, so the stream-ordered blocking copy is the synchronisation --
        # a device-wide sync would also wait for unrelated streams (e.g. the caller's H2D prefetch).
        t_wait = time.perf_counter() if HOST_WAIT is not None else 0.0
        rec = torch.cat([info.to(torch.float64), (flag if flag is not None else info.new_zeros(1)).to(torch.float64),
                         out["loss"]]).cpu()
        if HOST_WAIT is not None:
            HOST_WAIT[0] += time.perf_counter() - t_wait
        self._release()
        if rec[1] != 0:
            raise RuntimeError("GPTQ: a batch assumed identical for linears sharing one Hessian accumulator differed "
                               "(the tensor was modified between their forward hooks); set GPTQ_B200_SHARE_H=0")
        bad = int(rec[0])
        if bad != 0:
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: The factorization could not be completed because the input is not positive-definite "
                f"(the leading minor of order {bad} is not positive-definite).")
        self._error, self._error_dev = float(rec[2]), out["loss"]
        print('time %.2f' % (time.time() - tick))
        print('error', self.error)

        Wout = out["Wout"]
        if row_range is None:
            if isinstance(self.layer, _Conv1D):
                Wfull = Wout.t()
            else:
                Wfull = Wout
            self.layer.weight.data = Wfull.reshape(self.layer.weight.shape)
        self._maxq = qcfg[0]
        self.Wq = Wout                                  # [rows(shard), K] dequantised weights in the layer dtype
        self.codes = out["codes"]
        self.scales = out["scale"]
        self.zeros = out["zero"]
        self.perm = perm
        self.g_idx = out["g_idx"]
        # quantizer state the reference leaves behind: find_params on the full rows (gptq.py:181-185, quant.py:39,105-109) ...
        if out["full_scale"] is not None:
            q.maxq = torch.full((), qcfg[0], dtype=q.maxq.dtype, device=dev) if q.maxq.device != dev else q.maxq
            q.scale, q.zero = out["full_scale"], out["full_zero"]
        # ... then, with groups, the parameters of the group of the LAST processed column
        q_after = q
        if out["q_scale"] is not None:
            if static_groups:
                q_after = q.shallow_clone(copy_buffers=False)   # scale / zero are replaced right below; the reference swaps in groups[...] (a deepcopy), gptq.py:260
            q_after.scale, q_after.zero = out["q_scale"], out["q_zero"]
        self.quantizer = q_after
        if DEBUG:
            print(torch.sum((self.layer(self.inp1) - self.out1) ** 2))

    def free(self):
        if DEBUG:
            self.inp1 = None
            self.out1 = None
        self.H = None
        self.Losses = None
        self.Trace = None
        # The reference calls torch.cuda.empty_cache() here (gptq.py:318), which forces a cudaFree/cudaMalloc
        # round trip (and a device sync) for every linear; the caching allocator makes it unnecessary.
        if EMPTY_CACHE_ON_FREE:
            torch.cuda.empty_cache()
