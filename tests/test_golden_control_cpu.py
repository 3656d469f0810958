"""The control experiment behind the size-dependent rounding-tie tolerance of tests/test_gpu_headline.py: the committed
outputs of tools/control_chain_sensitivity.py (the oracle against itself with an fp64 chain) say what the tolerance comment
claims."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(k):
    with open(os.path.join(GOLDEN, f"control_chain_sensitivity_k{k}.json")) as f:
        return json.load(f)


def test_reference_reproduces_itself_at_k4096():
    d = _load(4096)
    assert d["tol_2^-18"]["rows_bad"] == 0


def test_reference_flips_near_ties_at_k11008():
    d = _load(11008)
    tight, loose = d["tol_2^-18"], d["tol_2^-16"]
    assert tight["rows_bad"] > 0 and tight["max_rounding_margin"] > 2.0 ** -18      # the tight tolerance does not cover the reference itself
    assert tight["max_rounding_margin"] < 2.0 ** -16 and loose["unexplained_rows"] == 0
