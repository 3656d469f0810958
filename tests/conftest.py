import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def cpu_id():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return ""


def same_cpu_as_golden():
    with open(os.path.join(GOLDEN, "PROVENANCE.json")) as f:
        return json.load(f)["cpu"] == cpu_id()


@pytest.fixture(scope="session")
def golden_A():
    return load_golden("gptq_A.npz")


@pytest.fixture(scope="session")
def golden_B():
    return load_golden("gptq_B.npz")


@pytest.fixture(scope="session")
def golden_quantizer():
    return load_golden("quantizer.npz")
