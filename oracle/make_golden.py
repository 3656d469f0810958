"""Generate tests/golden/*.npz by running the REFERENCE's own code (JerryZhou54/gptq).

Run in the build container only (needs /root/reference, read-only):

    python oracle/make_golden.py

It imports ``/root/reference/gptq.py`` and ``quant.py`` unmodified, with two shims
(SURVEY 8c): stub ``matplotlib`` modules (gptq.py:14 -> plot_activation.py:1) and a no-op
``torch.cuda.synchronize`` (gptq.py:292) because this container has no GPU.  Inputs are
seeded synthetic tensors; inputs AND reference outputs are stored so the fixtures do not
depend on any RNG implementation.  Nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import platform
import re
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    torch.cuda.synchronize = lambda *a, **k: None
    import gptq as ref_gptq          # noqa
    import quant as ref_quant        # noqa
    sys.path.pop(0)
    return ref_gptq, ref_quant


def cpu_id() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def make_inputs(seed, R, K, nbatch, tokens, rank3, dead=(), outliers=()):
    g = torch.Generator().manual_seed(seed)
    chan = torch.exp(0.5 * torch.randn(K, generator=g))
    X = []
    for _ in range(nbatch):
        x = torch.randn(tokens, K, generator=g) * chan
        for c in outliers:
            x[:, c] *= 20
        for c in dead:
            x[:, c] = 0
        x = x.half()
        X.append(x.unsqueeze(0) if rank3 else x)
    W = (torch.randn(R, K, generator=g) * 0.02).half()
    return X, W


CASES_A = [
    dict(name="b4_perchannel", bits=4),
    dict(name="b3_g128", bits=3, groupsize=128),
    dict(name="b4_g128_act_static", bits=4, groupsize=128, actorder=True, static_groups=True),
    dict(name="b2_g64", bits=2, groupsize=64),
    dict(name="b3_act", bits=3, actorder=True),
    dict(name="b4_g32_static", bits=4, groupsize=32, static_groups=True),
    dict(name="b4_g64_act", bits=4, groupsize=64, actorder=True),
    dict(name="b8_sym", bits=8, sym=True),
    dict(name="b4_mse", bits=4, mse=True),
    dict(name="trits", bits=2, trits=True),
    dict(name="b4_g1024", bits=4, groupsize=1024),
    dict(name="b3_perchannel", bits=3),
    dict(name="b2_perchannel", bits=2),
]
CASES_B = [
    dict(name="b4_perchannel", bits=4),
    dict(name="b3_g64_act_static", bits=3, groupsize=64, actorder=True, static_groups=True),
    dict(name="b4_g128", bits=4, groupsize=128),
    dict(name="b4_g128_static_pd1", bits=4, groupsize=128, static_groups=True, percdamp=0.1),
]


def run_set(ref_gptq, ref_quant, tag, seed, R, K, nbatch, tokens, rank3, cases, dead=(), outliers=()):
    X, W = make_inputs(seed, R, K, nbatch, tokens, rank3, dead, outliers)
    out = {"W": W.numpy(), "X": torch.stack(X).numpy()}
    meta = {"R": R, "K": K, "nbatch": nbatch, "tokens": tokens, "rank3": rank3, "cases": []}
    H_ref = None
    for case in cases:
        layer = nn.Linear(K, R, bias=False)
        layer.weight.data = W.clone()
        g = ref_gptq.GPTQ(layer)
        g.quantizer = ref_quant.Quantizer()
        g.quantizer.configure(case["bits"], perchannel=True, sym=case.get("sym", False),
                              mse=case.get("mse", False), trits=case.get("trits", False))
        for x in X:
            g.add_batch(x, None)
        if H_ref is None:
            H_ref = g.H.clone()
            out["H"] = H_ref.numpy()
            out["input_mean"] = g.input.numpy()
            meta["nsamples"] = int(g.nsamples)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
            g.fasterquant(blocksize=128, percdamp=case.get("percdamp", .01),
                          groupsize=case.get("groupsize", -1),
                          actorder=case.get("actorder", False),
                          static_groups=case.get("static_groups", False))
        err = float(re.search(r"error ([-+0-9.eE]+|nan|inf)", buf.getvalue()).group(1))
        n = case["name"]
        out[f"{n}/Q"] = layer.weight.data.numpy()                  # fp16, as the drivers see it
        out[f"{n}/scale"] = g.quantizer.scale.numpy()
        out[f"{n}/zero"] = g.quantizer.zero.numpy()
        c = dict(case)
        c["error"] = err
        meta["cases"].append(c)
        # 3-bit / 4-bit packing through the reference's own packers (per-row scales only)
        if case.get("groupsize", -1) == -1 and not case.get("trits", False) and K % 32 == 0:
            if case["bits"] == 3:
                ql = ref_quant.Quant3Linear(K, R)
                ql.pack(layer, g.quantizer.scale, g.quantizer.zero)
                out[f"{n}/qweight3"] = ql.qweight.numpy()
                out[f"{n}/pack_zeros"] = ql.zeros.numpy()
            if case["bits"] in (2, 4):
                # zeroShot/models/quant.py:172-185 (Quant4Linear.__init__), executed literally
                b = case["bits"]
                zs = g.quantizer.zero.clone() * g.quantizer.scale
                iw = torch.round((layer.weight.data + zs) / g.quantizer.scale).to(torch.int).t().contiguous()
                if b == 4:
                    qw = torch.zeros((iw.shape[0] // 8, iw.shape[1]), dtype=torch.int)
                    for i in range(iw.shape[0]):
                        qw[i // 8] |= iw[i] << (4 * (i % 8))
                    out[f"{n}/qweight4"] = qw.numpy()
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, f"gptq_{tag}.npz"), **out)
    print(tag, "cases:", [c["name"] for c in meta["cases"]])


def run_quantizer(ref_quant):
    """Quantizer.find_params / quantize over the layouts quant.py:37-118 handles."""
    g = torch.Generator().manual_seed(7)
    out, meta = {}, []
    W = torch.randn(16, 96, generator=g) * 0.05
    W[3] = 0                                       # all-zero row -> (-1,+1) (quant.py:65-67)
    W[5] = W[5].abs()                              # xmin clamps to 0
    out["W"] = W.numpy()
    A3 = torch.randn(2, 10, 24, generator=g)
    out["A3"] = A3.numpy()
    cfgs = [
        dict(name="w_b4", x="W", bits=4, perchannel=True, sym=False),
        dict(name="w_b4_sym", x="W", bits=4, perchannel=True, sym=True),
        dict(name="w_b3_mse", x="W", bits=3, perchannel=True, sym=False, mse=True),
        dict(name="w_b2_sym_mse", x="W", bits=2, perchannel=True, sym=True, mse=True),
        dict(name="w_trits", x="W", bits=2, perchannel=True, sym=False, trits=True),
        dict(name="w_b4_tensor", x="W", bits=4, perchannel=False, sym=False),
        dict(name="w_b8", x="W", bits=8, perchannel=True, sym=False),
        dict(name="a3_b4", x="A3", bits=4, perchannel=True, sym=False, weight=False),
        dict(name="a3_b4_tensor", x="A3", bits=4, perchannel=False, sym=True, weight=False),
    ]
    for c in cfgs:
        q = ref_quant.Quantizer()
        q.configure(c["bits"], perchannel=c["perchannel"], sym=c["sym"], mse=c.get("mse", False),
                    trits=c.get("trits", False))
        x = {"W": W, "A3": A3}[c["x"]]
        q.find_params(x.clone(), weight=c.get("weight", True))
        out[c["name"] + "/scale"] = q.scale.numpy()
        out[c["name"] + "/zero"] = q.zero.numpy()
        out[c["name"] + "/xq"] = q.quantize(x.clone()).numpy()
        meta.append(c)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "quantizer.npz"), **out)
    print("quantizer cases:", [c["name"] for c in meta])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)           # fixtures independent of the thread count
    ref_gptq, ref_quant = import_reference()
    run_set(ref_gptq, ref_quant, "A", seed=0, R=32, K=256, nbatch=3, tokens=128, rank3=True, cases=CASES_A)
    run_set(ref_gptq, ref_quant, "B", seed=1, R=24, K=200, nbatch=2, tokens=160, rank3=False, cases=CASES_B,
            dead=(17, 150), outliers=(3, 64, 65, 199))
    run_quantizer(ref_quant)
    with open(os.path.join(OUT, "PROVENANCE.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference": "JerryZhou54/gptq @ /root/reference",
                   "torch": torch.__version__, "numpy": np.__version__, "cpu": cpu_id(),
                   "threads": torch.get_num_threads()}, f, indent=1)


if __name__ == "__main__":
    main()
