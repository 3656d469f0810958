"""TEST INFRASTRUCTURE ONLY -- CPU restatements (numpy, fp64) of the two algorithms the CUDA chain uses instead of
the reference's three LAPACK calls (gptq.py:177-179: cholesky, cholesky_inverse, cholesky(upper=True)).  They exist so
that the mathematics of csrc/chol.cu can be checked on a machine without a GPU; nothing in gptq_b200/ imports them.

U is the unique upper-triangular matrix with positive diagonal and H^-1 = U^T U, i.e. U = R^-1 where H = R R^T with R
upper triangular ("backward" Cholesky).
"""
import numpy as np


def fused_block(A):
    """chol_diag3/4_kernel: backward Cholesky A = R R^T and V = R^-1 in ONE sweep k = n-1 .. 0.
    With T = diag(R) R^-1:  T[i,:] = e_i - sum_{k>i} R[i,k] V[k,:],  V[k,:] = T[k,:] / R[k,k].
    Step k: column k of R and row k of V are final; two rank-1 updates on complementary column ranges."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    T = np.eye(n)
    rinv = np.zeros(n)
    for k in range(n - 1, -1, -1):
        d = A[k, k]
        if not d > 0:
            raise np.linalg.LinAlgError(f"leading minor of order {k + 1} is not positive-definite")
        rinv[k] = 1.0 / np.sqrt(d)
        c = A[:k, k] * rinv[k]                      # R[i,k], i < k
        v = T[k, k:] * rinv[k]                      # V[k,c], c >= k
        A[:k, :k] -= np.outer(c, c)                 # Schur update of the factorisation (cols < k)
        T[:k, k:] -= np.outer(c, v)                 # elimination of the inverse        (cols >= k)
    return T * rinv[:, None]                        # V = diag(R)^-1 T


def hinv_upper_recursive(H, block=128, leaf=fused_block):
    """hinv_rec in csrc/chol.cu on block ranges [a,b):
        f(m,b);  Tt = V22 A12^T;  A11 -= Tt^T Tt;  f(a,m);  S = V11 Tt^T;  V12 = -S V22."""
    A = np.array(H, dtype=np.float64)
    n = A.shape[0]
    assert n % block == 0
    V = np.zeros_like(A)

    def f(a, b):
        if b - a == 1:
            s = slice(a * block, b * block)
            V[s, s] = leaf(A[s, s])
            return
        m = a + (b - a) // 2
        s1, s2 = slice(a * block, m * block), slice(m * block, b * block)
        f(m, b)
        Tt = V[s2, s2] @ A[s1, s2].T
        A[s1, s1] -= Tt.T @ Tt
        f(a, m)
        V[s1, s2] = -(V[s1, s1] @ Tt.T) @ V[s2, s2]
    f(0, n // block)
    return V


def syrk_tile_order(nbm, sw=12, sh=12):
    """Python restatement of decode_tile() in csrc/syrk_tc.cu: the order in which the Hessian kernel walks the
    128-row x 256-column tiles (bi, bj) with bi >= 2 bj -- 2-D chunks of sw column blocks x sh row blocks."""
    nbn = (nbm + 1) >> 1
    out = []
    for sc in range(0, nbn, sw):
        w = min(sw, nbn - sc)
        for r0 in range(2 * sc, nbm, sh):
            r1 = min(r0 + sh, nbm)
            for j in range(sc, sc + w):
                lo = max(r0, 2 * j)
                for bi in range(lo, r1):
                    out.append((bi, j))
    return out


def syrk2_tile_order(nb, sw=12, sh=6):
    """Python restatement of decode_tile2() in csrc/syrk_tc2.cu: the order in which the CTA-pair Hessian kernel walks the
    256 x 256 tiles (bi, bj) with bi >= bj -- chunks of sw column blocks x sh row blocks, column by column inside a chunk."""
    out = []
    for sc in range(0, nb, sw):
        w = min(sw, nb - sc)
        for r0 in range(sc, nb, sh):
            r1 = min(r0 + sh, nb)
            for j in range(sc, sc + w):
                for bi in range(max(r0, j), r1):
                    out.append((bi, j))
    return out


class Syrk2TileWalk:
    """Python restatement of TileWalk2 in csrc/syrk_tc2.cu: the same enumeration walked FORWARD (seek(t) with non-decreasing t),
    so that a cluster visiting tiles cid, cid + ncl, ... does not search from the start for every item."""

    def __init__(self, nb, sw=12, sh=6):
        self.nb, self.sw, self.sh = nb, sw, sh
        self.sc = self.r0 = self.j = self.base = 0

    def seek(self, t):
        while True:
            w = min(self.sw, self.nb - self.sc)
            r1 = min(self.r0 + self.sh, self.nb)
            lo = max(self.r0, self.j)
            cnt = r1 - lo
            if cnt > 0:
                if t < self.base + cnt:
                    return lo + (t - self.base), self.j
                self.base += cnt
            self.j += 1
            if self.j >= self.sc + w:
                self.r0 += self.sh
                self.j = self.sc
                if self.r0 >= self.nb:
                    self.sc += self.sw
                    self.r0 = self.j = self.sc
                    if self.sc >= self.nb:
                        return self.nb - 1, 0
