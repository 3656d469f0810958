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


class LookaheadSchedule:
    """hinv_lookahead in csrc/chol.cu restated as an explicit task list: the same operations, streams ("crit", "trail",
    "inv"), event records and waits, in the same issue order.  Each task carries the block-level regions it reads and writes
    (including the below-diagonal blocks a tile-granular "upper tiles" GEMM rewrites with garbage), so that the schedule
    can be checked for races without a GPU, and a `run` closure on fp64 numpy arrays, so that any order the events allow
    can be executed and compared.

    Panels of `pb` blocks, last to first; p = [e0, e1), next q = [f0, f1 = e0):
        crit : rec(p)  [wait trail(p+1)]  Tt_p[:, q] = V_pp A[q, p]^T   record rec(p)   A[q, q] -= Tt_p[:, q]^T Tt_p[:, q]
        trail: wait rec(p)   Tt_p[:, :f0] = V_pp A[:f0, p]^T   A[:f0, :f1] -= Tt_p[:, :f0]^T Tt_p[:, :f1]   record trail(p)
        inv  : wait rec(p)   V[p, e1:] = -V_pp S_p   S_q = R[q, e0:] V[e0:, e0:]   (S of the NEXT panel: it needs nothing of q)
    """

    def __init__(self, nb, pb=2, block=128, leaf=fused_block):
        self.nb, self.pb, self.block, self.leaf = nb, pb, block, leaf
        self.tasks = []
        np_ = (nb + pb - 1) // pb
        self.npanels = np_
        B = block

        def rect(mat, r0, r1, c0, c1):
            return {(mat, r, c) for r in range(r0, r1) for c in range(c0, c1)}

        def upper_rect(mat, r0, r1, c0, c1):
            # tiles of 1 x 2 blocks relative to the origin of C; a tile is written when it is not strictly below the diagonal
            return {(mat, r, c) for r in range(r0, r1) for c in range(c0, c1) if 2 * ((c - c0) // 2) + 1 >= (r - r0)}

        def add(stream, name, reads, writes, run, waits=(), record=None):
            self.tasks.append(dict(stream=stream, name=name, reads=reads, writes=writes, run=run, waits=tuple(waits), record=record))

        for i in range(np_ - 1, -1, -1):
            b1 = nb - (np_ - 1 - i) * pb
            b0 = max(b1 - pb, 0)
            last = i == np_ - 1
            P = slice(b0 * B, b1 * B)

            def rec(A, V, S, b0=b0, b1=b1, P=P):
                V[P, P] = hinv_upper_recursive(A[P, P], B, leaf) if (b1 - b0) > 1 else leaf(A[P, P])
                # the in-panel recursion leaves R12^T of the panel in the lower mirror of A[P, P]
                Rp = np.linalg.inv(V[P, P])
                A[P, P] = np.triu(A[P, P]) + np.tril(Rp.T, -1)
            add("crit", f"rec{i}", rect("A", b0, b1, b0, b1), rect("A", b0, b1, b0, b1) | rect("V", b0, b1, b0, b1) | {("Srec", 0, 0)}, rec)
            if i > 0:
                a0 = max(b0 - pb, 0)
                Q = slice(a0 * B, b0 * B)

                def solve_a(A, V, S, P=P, Q=Q):
                    A[P, Q] = V[P, P] @ A[Q, P].T
                add("crit", f"solveA{i}", rect("V", b0, b1, b0, b1) | rect("A", a0, b0, b0, b1), rect("A", b0, b1, a0, b0), solve_a,
                    waits=() if last else (f"trail{i + 1}",), record=f"rec{i}")

                def upd_a(A, V, S, P=P, Q=Q):
                    A[Q, Q] -= np.triu(A[P, Q].T @ A[P, Q])
                add("crit", f"updA{i}", rect("A", b0, b1, a0, b0) | rect("A", a0, b0, a0, b0), upper_rect("A", a0, b0, a0, b0), upd_a)
                if a0 > 0:
                    F = slice(0, a0 * B)
                    E = slice(0, b0 * B)

                    def solve_b(A, V, S, P=P, F=F):
                        A[P, F] = V[P, P] @ A[F, P].T
                    add("trail", f"solveB{i}", rect("V", b0, b1, b0, b1) | rect("A", 0, a0, b0, b1), rect("A", b0, b1, 0, a0), solve_b,
                        waits=(f"rec{i}",))

                    def trail_b(A, V, S, P=P, F=F, E=E):
                        upd = A[P, F].T @ A[P, E]
                        rows = np.arange(F.stop)[:, None]
                        cols = np.arange(E.stop)[None, :]
                        A[F, E] -= np.where(cols >= rows, upd, 0.0)
                    add("trail", f"trailB{i}", rect("A", b0, b1, 0, b0) | upper_rect("A", 0, a0, 0, b0), upper_rect("A", 0, a0, 0, b0), trail_b,
                        record=f"trail{i}")
                else:
                    add("trail", f"trailB{i}", set(), set(), lambda A, V, S: None, waits=(f"rec{i}",), record=f"trail{i}")
            else:
                add("crit", f"mark{i}", set(), set(), lambda A, V, S: None, record=f"rec{i}")
            # inverse stream: V[p, e1:] = -V_pp S_p with the S_p that was issued one panel earlier, then S of the next panel
            if not last:
                L = slice(b1 * B, nb * B)

                def inv_v(A, V, S, P=P, L=L):
                    V[P, L] = -V[P, P] @ S["inv"]
                add("inv", f"invV{i}", rect("V", b0, b1, b0, b1) | {("Sinv", 0, 0)}, rect("V", b0, b1, b1, nb), inv_v, waits=(f"rec{i}",))
            if i > 0:
                a0 = max(b0 - pb, 0)
                Q = slice(a0 * B, b0 * B)
                L2 = slice(b0 * B, nb * B)

                def inv_s(A, V, S, Q=Q, L2=L2):
                    S["inv"] = A[L2, Q].T @ V[L2, L2]
                add("inv", f"invS{i - 1}", rect("A", b0, nb, a0, b0) | rect("V", b0, nb, b0, nb), {("Sinv", 0, 0)}, inv_s,
                    waits=(f"rec{i}",) if last else ())

    def happens_before(self):
        """reach[i][j]: task i is ordered before task j by stream order and the events."""
        n = len(self.tasks)
        succ = [set() for _ in range(n)]
        prev_in_stream, recorded = {}, {}
        for j, t in enumerate(self.tasks):
            if t["stream"] in prev_in_stream:
                succ[prev_in_stream[t["stream"]]].add(j)
            prev_in_stream[t["stream"]] = j
            for w in t["waits"]:
                succ[recorded[w]].add(j)          # KeyError = a wait issued before its record (a deadlock on the device)
            if t["record"]:
                recorded[t["record"]] = j
        reach = [set() for _ in range(n)]
        for i in range(n - 1, -1, -1):
            for j in succ[i]:
                reach[i] |= {j} | reach[j]
        return succ, reach

    def races(self):
        _, reach = self.happens_before()
        bad = []
        for i, a in enumerate(self.tasks):
            for j in range(i + 1, len(self.tasks)):
                b = self.tasks[j]
                if j in reach[i]:
                    continue
                if (a["writes"] & (b["reads"] | b["writes"])) or (b["writes"] & a["reads"]):
                    bad.append((a["name"], b["name"]))
        return bad

    def run(self, H, order=None):
        """Execute the tasks in `order` (a permutation consistent with happens_before; default: issue order)."""
        A = np.array(H, dtype=np.float64)
        V = np.zeros_like(A)
        S = {}
        for k in (order if order is not None else range(len(self.tasks))):
            self.tasks[k]["run"](A, V, S)
        return V

    def random_order(self, rng):
        succ, _ = self.happens_before()
        n = len(self.tasks)
        indeg = [0] * n
        for i in range(n):
            for j in succ[i]:
                indeg[j] += 1
        ready = [i for i in range(n) if indeg[i] == 0]
        out = []
        while ready:
            k = ready.pop(int(rng.integers(len(ready))))
            out.append(k)
            for j in succ[k]:
                indeg[j] -= 1
                if indeg[j] == 0:
                    ready.append(j)
        assert len(out) == n
        return out


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
