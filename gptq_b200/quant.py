"""Host-side mirror of the reference's ``quant.py`` (JerryZhou54/gptq) over the sm_100a C ABI.

Same names, arguments and behaviour as the reference so ``from quant import *`` keeps working in
opt.py / llama.py / bloom.py / gemma.py:

* ``quantize(x, scale, zero, maxq)``                      -- quant.py:6-10
* ``Quantizer.configure / find_params / quantize / enabled / ready`` -- quant.py:12-131
* ``Quant3Linear(infeatures, outfeatures, faster).pack(linear, scales, zeros)`` -- quant.py:140-187
* ``make_quant3(module, names, name, faster)``            -- quant.py:205-216

All arithmetic runs in libgptq_b200.so on the GPU; there is no CPU path (CPU tensors handed to
``pack`` -- the reference drivers move layers to the CPU before packing, opt.py:362-373 -- are
copied to the device, packed there and the words are returned on the caller's device).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

__all__ = ["quantize", "Quantizer", "QuantLinear", "Quant2Linear", "Quant3Linear", "Quant4Linear",
           "make_quant3", "make_quant", "pack_codes", "unpack_codes"]


def _as_int(maxq) -> int:
    return int(maxq.item()) if torch.is_tensor(maxq) else int(maxq)


def _row_col_strides(x: torch.Tensor, p: torch.Tensor):
    """View ``x`` as [rows, cols] and express broadcasting of ``p`` as idx = r*s_row + c*s_col."""
    if p.numel() == 1:
        return x.reshape(1, -1) if x.dim() < 2 else x.reshape(-1, x.shape[-1]), p.reshape(1), 0, 0
    if p.dim() == x.dim() and all(d == 1 for d in p.shape[1:]) and p.shape[0] == x.shape[0]:
        return x.reshape(x.shape[0], -1), p.reshape(-1), 1, 0                 # [R,1,...] weight layout
    if p.dim() == x.dim() and all(d == 1 for d in p.shape[:-1]) and p.shape[-1] == x.shape[-1]:
        return x.reshape(-1, x.shape[-1]), p.reshape(-1), 0, 1                # [1,..,C] activation layout
    if x.dim() == 4 and p.dim() == 4 and p.shape[0] == 1 and p.shape[2] == 1 and p.shape[3] == 1:
        rows = p.expand(x.shape[0], x.shape[1], 1, 1).reshape(-1)             # [1,C,1,1] conv activations
        return x.reshape(x.shape[0] * x.shape[1], -1), rows, 1, 0
    raise NotImplementedError(f"quantize: unsupported broadcast of {tuple(p.shape)} against {tuple(x.shape)}")


def quantize(x, scale, zero, maxq):
    """quant.py:6-10 on the GPU (``maxq < 0`` = trits)."""
    _lib.require_cuda(x, "quantize: x")
    lib = _lib.load()
    maxq_i = _as_int(maxq)
    xf = x.float().contiguous()
    scale = scale.to(device=x.device, dtype=torch.float32)
    zero = zero.to(device=x.device, dtype=torch.float32)
    x2, s1, s_row, s_col = _row_col_strides(xf, scale)
    _, z1, _, _ = _row_col_strides(xf, zero.expand_as(scale) if zero.shape != scale.shape else zero)
    x2 = x2.contiguous()
    s1 = s1.contiguous()
    z1 = z1.contiguous()
    out = torch.empty_like(x2)
    R, Cc = x2.shape
    _lib.check(lib.gptq_b200_quantize(x2.data_ptr(), R, Cc, Cc, s1.data_ptr(), z1.data_ptr(), s_row, s_col, maxq_i,
                                      out.data_ptr(), Cc, _lib.stream()), "quantize")
    return out.reshape(x.shape)


def find_params_groups(W: torch.Tensor, maxq: int, sym: bool, mse: bool = False, norm: float = 2.4, grid: int = 100,
                       maxshrink: float = .8, groupsize: int = -1, col0: int = 0, ncols: int = None):
    """(scale, zero) [R, ngroups] for W[:, col0:col0+ncols] cut into groups of ``groupsize`` columns."""
    lib = _lib.load()
    assert W.dtype == torch.float32 and W.dim() == 2 and W.stride(1) == 1
    R, K = W.shape
    ncols = K - col0 if ncols is None else ncols
    gs = groupsize if groupsize > 0 else ncols
    ng = (ncols + gs - 1) // gs
    scale = torch.empty((R, ng), device=W.device, dtype=torch.float32)
    zero = torch.empty((R, ng), device=W.device, dtype=torch.float32)
    _lib.check(lib.gptq_b200_find_params(W.data_ptr(), R, K, W.stride(0), col0, ncols, groupsize, maxq, int(sym), int(mse),
                                         norm, grid, maxshrink, scale.data_ptr(), zero.data_ptr(), ng, 0, _lib.stream()),
               "find_params")
    return scale, zero


class Quantizer(nn.Module):
    """quant.py:12-131.  Buffers ``maxq``, ``scale``, ``zero`` so deepcopy / .cpu() / state_dict work."""

    def __init__(self, shape=1):
        super().__init__()
        self.register_buffer("maxq", torch.tensor(0))
        self.register_buffer("scale", torch.zeros(shape))
        self.register_buffer("zero", torch.zeros(shape))

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8, trits=False):
        self.wbits = bits
        self.maxq = torch.tensor(2 ** bits - 1)
        self.perchannel = perchannel
        self.sym = sym
        self.mse = mse
        self.norm = norm
        self.grid = grid
        self.maxshrink = maxshrink
        if trits:
            self.maxq = torch.tensor(-1)

    def find_params(self, x, weight=True):
        _lib.require_cuda(x, "Quantizer.find_params: x")
        dev = x.device
        maxq = _as_int(self.maxq)
        if self.maxq.device != dev:
            # quant.py:39 moves the buffer with .to(dev); a fill kernel does the same without a host-to-device copy, which
            # would queue behind whatever transfers the caller has in flight on other streams and stall this one
            self.maxq = torch.full((), maxq, dtype=self.maxq.dtype, device=dev)
        shape = x.shape
        if self.perchannel:                                 # quant.py:42-52
            if weight:
                x2 = x.flatten(1)
            else:
                if len(shape) == 4:
                    x2 = x.permute([1, 0, 2, 3]).flatten(1)
                elif len(shape) == 3:
                    x2 = x.reshape((-1, shape[-1])).t()
                else:
                    x2 = x.t()
        else:                                               # quant.py:53-54
            x2 = x.flatten().unsqueeze(0)
        x2 = x2.float().contiguous()
        scale, zero = find_params_groups(x2, maxq, self.sym, self.mse, self.norm, self.grid, self.maxshrink)
        scale = scale.reshape(-1)
        zero = zero.reshape(-1)
        if not self.perchannel:                             # quant.py:97-103
            tmp = shape[0] if weight else (shape[1] if len(shape) != 3 else shape[2])
            scale = scale.repeat(tmp)
            zero = zero.repeat(tmp)
        if weight:                                          # quant.py:105-109
            shp = [-1] + [1] * (len(shape) - 1)
            self.scale = scale.reshape(shp)
            self.zero = zero.reshape(shp)
            return
        if len(shape) == 4:
            self.scale = scale.reshape((1, -1, 1, 1))
            self.zero = zero.reshape((1, -1, 1, 1))
        if len(shape) == 3:
            self.scale = scale.reshape((1, 1, -1))
            self.zero = zero.reshape((1, 1, -1))
        if len(shape) == 2:
            self.scale = scale.unsqueeze(0)
            self.zero = zero.unsqueeze(0)

    def quantize(self, x):
        if self.ready():
            return quantize(x, self.scale, self.zero, self.maxq)
        return x

    def shallow_clone(self, copy_buffers=True):
        """A new Quantizer with the same configuration and buffers (what copy.deepcopy gives, without walking nn.Module)."""
        q = Quantizer()
        for k in ("wbits", "perchannel", "sym", "mse", "norm", "grid", "maxshrink", "trits"):
            if hasattr(self, k):
                setattr(q, k, getattr(self, k))
        if copy_buffers:
            q.maxq, q.scale, q.zero = self.maxq.clone(), self.scale.clone(), self.zero.clone()
        else:
            q.maxq, q.scale, q.zero = self.maxq, self.scale, self.zero
        return q

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)


# ------------------------------------------------------------------------------------------ packing
def pack_codes(codes: torch.Tensor, bits: int) -> torch.Tensor:
    """uint8 codes [R,K] on the GPU -> int32 qweight [K*bits/32, R] in the reference layout."""
    _lib.require_cuda(codes, "pack_codes: codes")
    assert codes.dtype == torch.uint8 and codes.dim() == 2 and codes.stride(1) == 1
    R, K = codes.shape
    if K % 32:
        raise ValueError("pack: in_features must be a multiple of 32")
    qweight = torch.empty((K * bits // 32, R), dtype=torch.int32, device=codes.device)
    _lib.check(_lib.load().gptq_b200_pack_codes(codes.data_ptr(), R, K, codes.stride(0), bits, qweight.data_ptr(),
                                                _lib.stream()), "pack_codes")
    return qweight


def unpack_codes(qweight: torch.Tensor, bits: int) -> torch.Tensor:
    _lib.require_cuda(qweight, "unpack_codes: qweight")
    assert qweight.dtype == torch.int32 and qweight.is_contiguous()
    K = qweight.shape[0] * 32 // bits
    R = qweight.shape[1]
    codes = torch.empty((R, K), dtype=torch.uint8, device=qweight.device)
    _lib.check(_lib.load().gptq_b200_unpack_codes(qweight.data_ptr(), R, K, bits, codes.data_ptr(), K, _lib.stream()),
               "unpack_codes")
    return codes


def _compute_device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise _lib.GptqB200Error("pack needs a CUDA device: gptq_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class QuantLinear(nn.Module):
    """Packed linear layer; ``bits`` in {2,3,4,8}.  ``Quant3Linear`` is the reference's class (bits=3)."""

    def __init__(self, bits, infeatures, outfeatures, faster=False, groupsize=-1):
        super().__init__()
        if infeatures % 32:
            raise ValueError("infeatures must be a multiple of 32")
        self.bits = bits
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.groupsize = groupsize
        ng = 1 if groupsize in (-1, None) else (infeatures + groupsize - 1) // groupsize
        # the reference's buffers (quant.py:144-149: zeros / scales [R,1], bias [R], qweight [K/32*bits, R]); with groups the
        # scale tables are [R, ngroups] and g_idx [K] maps every input channel to its group (act-order + static groups)
        self.register_buffer("zeros", torch.zeros((outfeatures, ng)))
        self.register_buffer("scales", torch.zeros((outfeatures, ng)))
        self.register_buffer("bias", torch.zeros(outfeatures))
        self.register_buffer("qweight", torch.zeros((infeatures // 32 * bits, outfeatures), dtype=torch.int))
        self.register_buffer("g_idx", torch.zeros(infeatures, dtype=torch.int32))
        self.faster = faster

    # ---- checkpoint format (opt.py:362-402: torch.save(model.state_dict()) / load_state_dict on a make_quant'ed model).
    # Per-row layers are saved with exactly the reference's four keys (no g_idx), so reference checkpoints load here and
    # per-row checkpoints written here load in the reference; grouped layers add their [R,ngroups] tables and g_idx.
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if self.scales.dim() < 2 or self.scales.shape[1] == 1:
            destination.pop(prefix + "g_idx", None)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for name in ("scales", "zeros"):
            t = state_dict.get(prefix + name)
            if t is not None and t.dim() == 2 and t.shape[0] == self.outfeatures and t.shape != getattr(self, name).shape:
                setattr(self, name, torch.zeros(t.shape, dtype=getattr(self, name).dtype, device=getattr(self, name).device))
        if prefix + "g_idx" not in state_dict:          # per-row checkpoint (the reference's format)
            state_dict = dict(state_dict)
            state_dict[prefix + "g_idx"] = torch.zeros(self.infeatures, dtype=torch.int32)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def pack(self, linear, scales, zeros, g_idx=None):
        """quant.py:152-187.  ``scales``/``zeros`` are [R,1] like the reference, or [R,ngroups] with ``g_idx`` [K]."""
        orig = linear.weight.device
        dev = _compute_device(linear.weight)
        scales = scales.to(dev).float()
        zeros = zeros.to(dev).float()
        self.zeros = (zeros * scales).to(orig)               # quant.py:153
        self.scales = scales.clone().to(orig)                 # quant.py:154
        if linear.bias is not None:
            self.bias = linear.bias.clone()
        W = linear.weight.data.to(dev)
        if W.dim() != 2:
            W = W.flatten(1)
        W = W.contiguous()
        R, K = W.shape
        s2 = scales.reshape(R, -1).contiguous()
        z2 = zeros.reshape(R, -1).contiguous()
        ng = s2.shape[1]
        gi = None
        if ng > 1:
            if g_idx is None:
                raise ValueError("grouped scales need g_idx")
            gi = g_idx.to(device=dev, dtype=torch.int32).contiguous()
        self.g_idx = gi.to(orig) if gi is not None else torch.zeros(K, dtype=torch.int32, device=orig)
        qweight = torch.empty((K * self.bits // 32, R), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gptq_b200_pack_weights(W.data_ptr(), _lib.dt(W), R, K, W.stride(0), s2.data_ptr(),
                                                          z2.data_ptr(), ng, _lib.ptr(gi), self.bits, qweight.data_ptr(),
                                                          _lib.stream()), "pack_weights")
        self.qweight = qweight.to(orig)

    def forward(self, x):
        """quant.py:189-203: single-token matvec on the packed weights (y starts as the bias)."""
        if x.shape[-1] == x.numel():
            _lib.require_cuda(x, "QuantLinear.forward: x")
            dev = x.device
            outshape = list(x.shape)
            outshape[-1] = self.bias.numel()
            dtype = x.dtype
            y = self.bias.to(device=dev, dtype=torch.float32).clone()
            xv = x.reshape(-1).contiguous()
            if xv.dtype not in (torch.float16, torch.bfloat16, torch.float32):
                xv = xv.float()
            qw = self.qweight.to(dev)
            s2 = self.scales.to(dev).float().reshape(self.outfeatures, -1).contiguous()
            z2 = self.zeros.to(dev).float().reshape(self.outfeatures, -1).contiguous()
            ng = s2.shape[1]
            gi = self.g_idx.to(device=dev, dtype=torch.int32).contiguous() if ng > 1 else None
            _lib.check(_lib.load().gptq_b200_vecquant_matvec(xv.data_ptr(), _lib.dt(xv), qw.data_ptr(), self.outfeatures,
                                                             self.infeatures, self.bits, s2.data_ptr(), z2.data_ptr(), ng,
                                                             _lib.ptr(gi), 0, y.data_ptr(), _lib.stream()), "vecquant_matvec")
            return y.to(dtype).reshape(outshape)
        raise ValueError('Only supports a single token currently.')


class Quant3Linear(QuantLinear):
    def __init__(self, infeatures, outfeatures, faster=False, groupsize=-1):
        super().__init__(3, infeatures, outfeatures, faster=faster, groupsize=groupsize)


class Quant4Linear(QuantLinear):
    """zeroShot/models/quant.py:162-200 (4-bit layout: code i of a column at bits 4*(i%8) of word i//8)."""

    def __init__(self, infeatures, outfeatures, faster=False, groupsize=-1):
        super().__init__(4, infeatures, outfeatures, faster=faster, groupsize=groupsize)


class Quant2Linear(QuantLinear):
    """2-bit layout by analogy with the 4-bit one (code << 2*(i%16)).  The reference defines NO 2-bit packed format, so this
    layout is PARITY-UNPINNED by construction; only the integer codes are checked against the oracle."""

    def __init__(self, infeatures, outfeatures, faster=False, groupsize=-1):
        super().__init__(2, infeatures, outfeatures, faster=faster, groupsize=groupsize)


def make_quant(module, names, bits, name="", faster=False, groupsize=-1):
    if isinstance(module, QuantLinear):
        return
    for attr in dir(module):
        tmp = getattr(module, attr)
        name1 = name + "." + attr if name != "" else attr
        if name1 in names:
            cls = {2: Quant2Linear, 3: Quant3Linear, 4: Quant4Linear}.get(bits)
            setattr(module, attr, cls(tmp.in_features, tmp.out_features, faster=faster, groupsize=groupsize) if cls
                    else QuantLinear(bits, tmp.in_features, tmp.out_features, faster=faster, groupsize=groupsize))
    for name1, child in module.named_children():
        make_quant(child, names, bits, name + "." + name1 if name != "" else name1, faster=faster, groupsize=groupsize)


def make_quant3(module, names, name="", faster=False):
    """quant.py:205-216."""
    make_quant(module, names, 3, name=name, faster=faster)


def find_linears(module, kinds=(nn.Conv2d, nn.Linear), name=""):
    """modelutils.py:8-16."""
    if type(module) in kinds or (QuantLinear in kinds and isinstance(module, QuantLinear)):
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_linears(child, kinds, name + "." + name1 if name != "" else name1))
    return res


def pack_model(model, quantizers, bits, extras=None, groupsize=-1, faster=False):
    """opt.py:362-373 (opt_pack3) for any of the packed widths: replace every quantized linear by its packed layer.
    `quantizers`: {name: Quantizer} as returned by the `*_sequential` driver.  The reference hands `pack` the quantizer's
    [R,1] scale / zero, which is only meaningful for groupsize -1 (with groups it is the LAST group's pair);
    grouped models therefore pack from `extras[name]` = {"scales" [R,ng], "zeros" [R,ng], "g_idx" [K]} (the attributes
    GPTQ.fasterquant leaves behind)."""
    layers = find_linears(model)
    layers = {n: layers[n] for n in quantizers}
    make_quant(model, quantizers, bits, faster=faster, groupsize=groupsize)
    qlayers = find_linears(model, (QuantLinear,))
    for name in qlayers:
        if name not in layers:
            continue
        ex = (extras or {}).get(name)
        if ex is not None and ex.get("scales") is not None and ex["scales"].dim() == 2 and ex["scales"].shape[1] > 1:
            qlayers[name].pack(layers[name], ex["scales"], ex["zeros"], g_idx=ex["g_idx"])
        else:
            q = quantizers[name]
            qlayers[name].pack(layers[name], q.scale, q.zero)
    return model


def load_quant(model, checkpoint, bits, groupsize=-1, skip=("lm_head", "model.decoder.project_out", "model.decoder.project_in"),
               faster=False):
    """opt.py:375-402 (load_quant3) without the model construction: `model` is a freshly built (unquantized, any init)
    model; its linears become packed layers and the checkpoint (a path or a state dict) is loaded strictly."""
    layers = find_linears(model)
    for name in skip:
        layers.pop(name, None)
    make_quant(model, layers, bits, faster=faster, groupsize=groupsize)
    sd = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "read") else checkpoint
    model.load_state_dict(sd)
    return model
