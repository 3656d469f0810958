"""Control experiment for the rounding-tie tolerance of the headline parity tests (tests/test_gpu_headline.py TIE_TOL):
how reproducible are the REFERENCE's own codes under an fp32-level change of its own arithmetic?  The oracle
(oracle/gptq_oracle.py, the reference's algorithm on the CPU) is run twice on the same W and H -- once as the reference
does it (three fp32 LAPACK calls), once with the upper factor computed in fp64 and rounded to fp32 -- and the rows whose
codes differ are classified by the rounding margin of their first differing column.  CPU only.

    python tools/control_chain_sensitivity.py [K] > tests/golden/control_chain_sensitivity_k<K>.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import gptq_oracle as O  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 11008
R, T, ns = 512, 2048, 4
g = torch.Generator().manual_seed(7)
W16 = (torch.randn(R, K, generator=g) * 0.02).half()
chan = torch.exp(0.5 * torch.randn(K, generator=g))
st = O.HessianState(K, H=torch.zeros((K, K)))
for j in range(ns):
    O.add_batch(st, (torch.randn(T, K, generator=g) * chan).half())
kw = dict(blocksize=128, groupsize=128, actorder=True, static_groups=True, stable_perm=True)
t0 = time.time()
ref = O.fasterquant(W16, st.H, O.QParams(bits=4), keep_margin=True, **kw)
ctl = O.fasterquant(W16, st.H, O.QParams(bits=4), chain_fp64=True, **kw)
out = {"what": "oracle (fp32 chain, as the reference) vs oracle (fp64 chain rounded to fp32), same W and H",
       "R": R, "K": K, "samples": ns, "tokens": T, "bits": 4, "groupsize": 128, "actorder": True, "static_groups": True,
       "seconds": round(time.time() - t0, 1)}
for name, tol in (("tol_2^-18", 2.0 ** -18), ("tol_2^-16", 2.0 ** -16)):
    out[name] = O.classify_mismatches(ctl.codes, ref, tol)
print(json.dumps(out, indent=1))
