"""Drop-in for the reference's top-level ``quant`` module (``from quant import *`` in opt.py / llama.py / ...)."""
from gptq_b200.quant import *         # noqa: F401,F403
from gptq_b200.quant import quantize, Quantizer, Quant3Linear, make_quant3  # noqa: F401
