"""One traced call of the look-ahead chain (GPTQ_B200_CHAIN_TRACE=1 prints the per-panel timeline to stderr).

    GPTQ_B200_CHAIN_TRACE=1 python tools/chain_trace.py K panel
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GPTQ_B200_CHAIN_TRACE", "0")
import torch  # noqa: E402
import tools.chain_lab as L  # noqa: E402

K, panel = int(sys.argv[1]), int(sys.argv[2])
r, _ = L.run(K, 1, panel, iters=1)
print(r)
