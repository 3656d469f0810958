"""Sweep the panel width of the look-ahead chain:  python tools/chain_sweep.py K panel [panel ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.chain_lab as L  # noqa: E402

K = int(sys.argv[1])
base, U0 = L.run(K, 0, 2, iters=3)
print(json.dumps(base), flush=True)
for panel in [int(a) for a in sys.argv[2:]]:
    r, U1 = L.run(K, 1, panel, iters=3)
    r["rel_diff_vs_recursion"] = float((U1 - U0).norm() / U0.norm())
    print(json.dumps(r), flush=True)
    del U1
