#!/usr/bin/env python3
"""ALS half-step timing probe (BASELINE configs[3] shape: 480 189 users x 17 770 items, d=256, Zipf item popularity).
usage: python profiles/als_probe.py [scale=0.1] [d=256] [seg=4096] [mean_pos=208] [factor=0]
factor: tkr_debug_set_als_factor (0 default, 1 blocked rounds at every width, 2 per-column / four-column loops)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "top-k-rec_b200"))
import topkrec  # noqa: E402


def synth(n_users, n_items, mean_pos, seed=0):
    rng = np.random.default_rng(seed)
    cnt = np.maximum(1, rng.poisson(mean_pos, n_users))
    pop = 1.0 / np.arange(1, n_items + 1); cdf = np.cumsum(pop / pop.sum())
    items = np.searchsorted(cdf, rng.random(int(cnt.sum()))).astype(np.int64).clip(0, n_items - 1)
    users = np.repeat(np.arange(n_users, dtype=np.int64), cnt)
    key = np.unique(users * n_items + items)                       # dedup per user
    return key // n_items, key % n_items


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    seg = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    mean_pos = int(sys.argv[4]) if len(sys.argv) > 4 else 208
    topkrec.lib().tkr_debug_set_als_factor(int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    n_users, n_items = int(480189 * scale), 17770
    t0 = time.time()
    users, items = synth(n_users, n_items, mean_pos)
    nnz = users.size
    u_ptr = np.zeros(n_users + 1, np.int64); np.cumsum(np.bincount(users, minlength=n_users), out=u_ptr[1:])
    by_i = np.argsort(items, kind="stable")
    i_ptr = np.zeros(n_items + 1, np.int64); np.cumsum(np.bincount(items, minlength=n_items), out=i_ptr[1:])
    us = topkrec.AlsSide(u_ptr, items.astype(np.int32), seg)
    its = topkrec.AlsSide(i_ptr, users[by_i].astype(np.int32), seg)
    print("synth %.1fs: users %d items %d nnz %d; user segs %d, item segs %d (%d split rows, %d slots, %.2f GB partial)" % (
        time.time() - t0, n_users, n_items, nnz, us.plan.n_segs, its.plan.n_segs, its.plan.n_multi, its.n_slots,
        topkrec.lib().tkr_als_partial_bytes(d, its.n_slots) / 1e9), flush=True)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    U = torch.rand(n_users, d, device="cuda", generator=g)
    V = torch.rand(n_items, d, device="cuda", generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for rep in range(3):
        ev[0].record()
        XX = topkrec.als_gram(V, its.rated_dev, 0.01, 0.01)
        ev[1].record()
        lu = topkrec.als_solve_rows(us, V, U, XX, 1.0, 0.01, 0.0, 0.01)
        ev[2].record()
        XXv = topkrec.als_gram(U, us.rated_dev, 0.01, 0.0)
        ev[3].record()
        li = topkrec.als_solve_rows(its, U, V, XXv, 1.0, 0.01, 10.0, 10.0, item_loss=True)
        ev[4].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        fl_gram = 2.0 * nnz * d * d
        fl_u = fl_gram + n_users * (2.0 / 3.0) * d ** 3
        fl_i = fl_gram + n_items * (2.0 / 3.0) * d ** 3
        print("rep %d: gramV %.3f ms | user step %.2f ms (%.1f TFLOP/s ref-algorithm, %.2f M rows/s) | gramU %.3f ms | item step %.2f ms (%.1f TFLOP/s) | loss %.6g" % (
            rep, t[0], t[1], fl_u / t[1] / 1e9, n_users / t[1] / 1e3, t[2], t[3], fl_i / t[3] / 1e9, float(lu.sum()) + float(li.sum())), flush=True)


if __name__ == "__main__":
    main()
