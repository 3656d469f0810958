"""Drop-in for the reference's top-level ``gptq`` module (``from gptq import *`` in opt.py / llama.py / ...)."""
from gptq_b200.gptq import *          # noqa: F401,F403
from gptq_b200.gptq import GPTQ, DEBUG  # noqa: F401
from gptq_b200.quant import *         # noqa: F401,F403  (the reference's gptq.py re-exports quant.*)
