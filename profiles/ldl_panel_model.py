#!/usr/bin/env python3
"""CPU model of the planned d=256 factorisation (DESIGN.md section 7, K5 plan (a)): the normal matrix lives in a packed,
16-byte-aligned column layout in shared memory and is factored A = M D^-1 M^T in 32-column panels -- each panel in 4-column
rounds (4x4 diagonal block factored by one thread, the panel's other rows eliminated against it, rank-4 update of the rest
of the PANEL), then one rank-32 update of the trailing matrix -- with the right-hand side riding along and a back substitution
over the stored raw columns.  Checks the index algebra (layout, masks, update ranges) and the fp32 accuracy against
np.linalg.solve; the thread mapping of the CUDA kernel is NOT modelled.
usage: python profiles/ldl_panel_model.py"""
import numpy as np

DP, PW, RW = 256, 32, 4


def col_origin(k, dp=DP):
    """v(k): offset of column k's virtual row 0; column k occupies [v(k)+k, v(k)+dp); v(k) % 4 == 0 (als_solve.cu experiment)."""
    r = k & 3
    return k * dp - k * (k + 1) // 2 + 6 * (k >> 2) + r * (r + 1) // 2


def check_layout():
    end = 0
    for k in range(DP):
        v = col_origin(k)
        assert v % 4 == 0 and v + k >= end, (k, v, end)
        end = v + DP
    return end                                                     # floats needed


def factor_solve(A, b):
    n = A.shape[0]
    S = np.zeros(check_layout(), np.float32)                       # the shared-memory image: lower triangle only
    for c in range(n):
        S[col_origin(c) + c: col_origin(c) + n] = A[c:, c]
    col = lambda c: S[col_origin(c): col_origin(c) + n]            # noqa: E731  (rows < c of this view belong to other columns)
    pinv = np.zeros(n, np.float32)
    z = b.astype(np.float32).copy()
    for c0 in range(0, n, PW):
        # ---- panel c0 .. c0+PW-1, rows c0 .. n-1, in rounds of RW columns
        for j0 in range(c0, c0 + PW, RW):
            # A: the 4x4 diagonal block (one thread)
            d44 = np.array([[col(j0 + q)[j0 + r] if r >= q else 0 for q in range(RW)] for r in range(RW)], np.float32)
            l = np.zeros((RW, RW), np.float32)
            for q in range(RW):
                pinv[j0 + q] = np.float32(1) / d44[q, q]
                for r in range(q + 1, RW):
                    l[r, q] = d44[r, q] * pinv[j0 + q]
                    for q2 in range(q + 1, r + 1):
                        d44[r, q2] -= l[r, q] * d44[q2, q]
            for q in range(RW):                                     # diagonal rows of the published raw columns
                for r in range(q, RW):
                    col(j0 + q)[j0 + r] = d44[r, q]
            # right-hand side inside the block, then B: rows below the block against it (raw M, in place)
            for q in range(RW):
                for q2 in range(q):
                    z[j0 + q] -= l[q, q2] * z[j0 + q2]
            rows = slice(j0 + RW, n)
            for q in range(RW):
                for q2 in range(q):
                    col(j0 + q)[rows] -= col(j0 + q2)[rows] * l[q, q2]
            # C: rank-4 update of the rest of the panel (columns j0+RW .. c0+PW-1, rows >= column) and of the right-hand side
            for q in range(RW):
                lq = col(j0 + q)[rows] * pinv[j0 + q]               # scaled L
                z[rows] -= lq * z[j0 + q]
                for c in range(j0 + RW, c0 + PW):
                    col(c)[c:n] -= (col(j0 + q)[c:n] * pinv[j0 + q]) * col(j0 + q)[c]
        # ---- trailing matrix: rank-PW update from the finished panel (every 8x8 tile: load, 32 k-steps, store)
        for c in range(c0 + PW, n):
            for k in range(c0, c0 + PW):
                col(c)[c:n] -= (col(k)[c:n] * pinv[k]) * col(k)[c]
    # back substitution: x_r = (z_r - sum_{q > r} M[q][r] x_q) / D_r
    x = np.zeros(n, np.float32)
    t = z.copy()
    for r in range(n - 1, -1, -1):
        x[r] = t[r] * pinv[r]
        t[:r] -= np.array([col(c)[r] for c in range(r)], np.float32) * x[r]
    return x


def main():
    print("layout: %d floats (%.1f KB) for DP=%d" % (check_layout(), check_layout() * 4 / 1024, DP))
    rng = np.random.default_rng(0)
    for name, gen in (("uniform(0,1) factors", lambda s: rng.random(s)), ("N(0,0.3) factors", lambda s: 0.3 * rng.standard_normal(s))):
        V = gen((4000, DP)).astype(np.float32)
        Vi = V[rng.integers(0, 4000, 208)]
        A = (0.01 * V.T @ V + 0.01 * np.eye(DP, dtype=np.float32) + 0.99 * Vi.T @ Vi).astype(np.float32)
        b = Vi.sum(0)
        x64 = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        x = factor_solve(A, b)
        print("%-22s cond %.1e  max|x - x64| / max|x64| = %.2e" % (name, np.linalg.cond(A.astype(np.float64)), np.abs(x - x64).max() / np.abs(x64).max()))
        assert np.abs(x - x64).max() / np.abs(x64).max() < 1e-4


if __name__ == "__main__":
    main()
