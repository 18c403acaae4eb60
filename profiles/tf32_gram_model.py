#!/usr/bin/env python3
"""CPU model for the planned tensor-core Gram (DESIGN.md section 7, K5 plan (b)): how far does a TF32 Gram matrix -- plain,
and as the 3-term split hi*hi + hi*lo + lo*hi with fp32 accumulation -- move the ALS solution, compared with the fp32 Gram
the kernel forms today and with the reference's fp32-BLAS Gram (all solved in fp64, so only the Gram differs)?
TF32 = fp32 with the mantissa rounded to 10 bits (round-to-nearest-even on the dropped 13 bits)."""
import numpy as np


def tf32(x):
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def gram_variants(Y):
    hi = tf32(Y); lo = tf32(Y - hi)
    g32 = Y.T @ Y
    g_tf = hi.T @ hi
    g_3x = hi.T @ hi + (hi.T @ lo + lo.T @ hi)
    return g32, g_tf.astype(np.float32), g_3x.astype(np.float32)


def main():
    rng = np.random.default_rng(0)
    for name, d, n_items, npos, gen in (("uniform(0,1) start, d=256, 5000 positives", 256, 17770, 5000, lambda s: rng.random(s)),
                                        ("uniform(0,1) start, d=256, 208 positives", 256, 17770, 208, lambda s: rng.random(s)),
                                        ("centred factors N(0,0.3), d=256, 208 positives", 256, 17770, 208, lambda s: 0.3 * rng.standard_normal(s)),
                                        ("uniform(0,1) start, d=64, 300 positives of 140 items", 64, 140, 300, lambda s: rng.random(s))):
        V = gen((n_items, d)).astype(np.float32)
        XX = (0.01 * (V.astype(np.float64).T @ V.astype(np.float64)) + 0.01 * np.eye(d))
        pos = rng.integers(0, n_items, npos)
        Vi = V[pos]
        rhs = Vi.astype(np.float64).sum(0)
        exact = np.linalg.solve(XX + 0.99 * (Vi.astype(np.float64).T @ Vi.astype(np.float64)), rhs)
        g32, g_tf, g_3x = gram_variants(Vi)
        perm = rng.permutation(npos)
        g32b = np.zeros((d, d), np.float32)
        for r in Vi[perm]:
            g32b += np.outer(r, r)                                      # fp32, another summation order
        out = {}
        for k, g in (("fp32 BLAS", g32), ("fp32 other order", g32b), ("tf32", g_tf), ("tf32 x3", g_3x)):
            x = np.linalg.solve(XX + 0.99 * g.astype(np.float64), rhs)
            out[k] = float(np.abs(x - exact).max() / np.abs(exact).max())
        print("%-52s cond %.1e  " % (name, np.linalg.cond(XX + 0.99 * (Vi.astype(np.float64).T @ Vi.astype(np.float64)))) +
              "  ".join("%s %.1e" % kv for kv in out.items()))


if __name__ == "__main__":
    main()
