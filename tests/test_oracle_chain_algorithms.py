"""CPU check of the mathematics behind csrc/chol.cu: the recursion and the fused factor + inverse diagonal block must
give the reference's upper factor (gptq.py:177-179)."""
import numpy as np
import pytest
import torch

from oracle import chain_algorithms as C


def _spd(n, seed):
    g = np.random.default_rng(seed)
    X = g.standard_normal((3 * n, n)) * np.exp(0.5 * g.standard_normal(n))
    return X.T @ X / n + 0.01 * np.eye(n)


def _reference_upper(H):
    Ht = torch.from_numpy(H)
    L = torch.linalg.cholesky(Ht)                                    # gptq.py:177
    return torch.linalg.cholesky(torch.cholesky_inverse(L), upper=True).numpy()   # gptq.py:178-179


@pytest.mark.parametrize("n", [1, 7, 32, 128])
def test_fused_block_equals_reference_route(n):
    H = _spd(n, n)
    V = C.fused_block(H)
    U = _reference_upper(H)
    assert np.allclose(np.tril(V, -1), 0)
    assert np.linalg.norm(V - U) <= 1e-10 * np.linalg.norm(U)
    assert np.allclose(V.T @ V @ H, np.eye(n), atol=1e-8)            # H^-1 = U^T U


@pytest.mark.parametrize("nb,block", [(1, 16), (2, 16), (5, 16), (8, 8), (3, 128)])
def test_recursion_equals_reference_route(nb, block):
    H = _spd(nb * block, nb)
    V = C.hinv_upper_recursive(H, block)
    U = _reference_upper(H)
    assert np.allclose(np.tril(V, -1), 0)
    assert np.linalg.norm(V - U) <= 1e-9 * np.linalg.norm(U)


def test_fused_block_reports_non_pd():
    H = _spd(16, 3)
    H[5, 5] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        C.fused_block(H)
