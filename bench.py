#!/usr/bin/env python3
"""Benchmark of the two hot paths on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (BASELINE.json configs[1]): BPR synthetic MovieLens-10M shape, 70k users x 10k
items, d=128, reference optimiser (RMSProp), one "step" = one synchronous mini-batch of
--batch triples (default 2^20) per GPU.  `value` = triples/s with the triples resident in
HBM; `e2e` = the same through tkr_bpr_step_host (triples in pinned host memory, H2D + loss
D2H inside the timed region).  The second path (score + top-30, BASELINE configs[4] slice:
18944 users x 1M items, d=128) is reported in the same line under "score_topk".
`--impl reference` times the CPU port of the reference step (oracle/bpr_ref.c, OpenMP, all
host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "top-k-rec_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

N_USERS, N_ITEMS, D = 70000, 10000, 128
METRIC = "bpr_triples_per_sec"


def algorithmic_bytes_per_triple(d, optimizer="rmsprop"):
    """read+write 3 param rows and 3 rms rows, 2 biases + 2 bias-rms, 3 ids."""
    return 48 * d + 44 if optimizer == "rmsprop" else 24 * d + 28


# ----------------------------------------------------------------------------- synthetic data
def synth_interactions(n_users=N_USERS, n_items=N_ITEMS, n_pairs=10_000_000, seed=0):
    """user ~ uniform, item ~ Zipf(1.0), like w.p. 0.15, dedup per user,
    every user >= 1 positive.  Returns (tr_users, pos_indptr, pos_idx) with items ascending."""
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, n_items + 1); pop /= pop.sum()
    users = rng.integers(0, n_users, n_pairs)
    items = rng.choice(n_items, n_pairs, p=pop)
    like = rng.random(n_pairs) < 0.15
    keys = np.unique(users[like] * n_items + items[like])
    have = np.zeros(n_users, bool); have[keys // n_items] = True
    missing = np.nonzero(~have)[0]
    keys = np.unique(np.concatenate([keys, missing * n_items + rng.choice(n_items, missing.size, p=pop)]))
    u, it = keys // n_items, (keys % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, np.int64); np.cumsum(np.bincount(u, minlength=n_users), out=indptr[1:])
    return np.arange(n_users, dtype=np.int32), indptr, it


def init_state_np(n_users, n_items, d, seed=1):
    rng = np.random.default_rng(seed)
    st = {"U": (0.01 * rng.standard_normal((n_users, d))).astype(np.float32),
          "V": (0.01 * rng.standard_normal((n_items, d))).astype(np.float32),
          "b": np.zeros(n_items, np.float32)}
    for k in ("U", "V", "b"):
        st["ms" + k] = np.ones_like(st[k])
    return st


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during a timed region (pynvml, 50 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._th = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True); self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th is not None:
            self._th.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def profile_traffic(kernel):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU port (reference arm / cpu_baseline)
def cpu_bpr_steps(batch, n_steps, warmup, seed=123):
    """Time the OpenMP C port of the reference step (oracle/bpr_ref.c) on the C2 workload.
    Returns (triples_per_sec, ms_per_step, threads)."""
    import ctypes
    from oracle import clib

    class Cfg(ctypes.Structure):
        _fields_ = [("n_users", ctypes.c_int32), ("n_items", ctypes.c_int32), ("d", ctypes.c_int32),
                    ("lu", ctypes.c_float), ("li", ctypes.c_float), ("lj", ctypes.c_float), ("lb", ctypes.c_float),
                    ("lr", ctypes.c_float), ("l1", ctypes.c_int32), ("sgd", ctypes.c_int32)]
    lib = clib.lib()
    st = init_state_np(N_USERS, N_ITEMS, D)
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, N_ITEMS + 1); pop /= pop.sum()
    u = rng.integers(0, N_USERS, batch).astype(np.int32)
    i = rng.choice(N_ITEMS, batch, p=pop).astype(np.int32)
    j = rng.integers(0, N_ITEMS, batch).astype(np.int32)
    c = Cfg(N_USERS, N_ITEMS, D, 2.5e-3, 2.5e-3, 2.5e-4, 0.0, 1e-4, 0, 0)
    loss = ctypes.c_double()
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

    def step():
        rc = lib.tkr_ref_bpr_step(ctypes.byref(c), fp(st["U"]), fp(st["V"]), fp(st["b"]), fp(st["msU"]), fp(st["msV"]),
                                  fp(st["msb"]), fp(u), fp(i), fp(j), ctypes.c_int64(batch), ctypes.byref(loss))
        assert rc == 0
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    dt = time.perf_counter() - t0
    return batch * n_steps / dt, 1e3 * dt / n_steps, os.cpu_count()


def run_reference(args, rank):
    if rank != 0:
        return
    batch = min(args.batch, 1 << 20)
    val, ms, threads = cpu_bpr_steps(batch, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "triples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BPR synthetic MovieLens-10M shape (70k users x 10k items), d=128, RMSProp, batch %d" % batch,
                       "n_users": N_USERS, "n_items": N_ITEMS, "d": D, "batch_size": batch},
            "cpu_baseline": {"value": val, "unit": "triples/s", "cores": threads, "kind": "port",
                             "sample": "%d steps of %d triples; OpenMP C port of single/bpr.py:71-101 + TF-1.15 RMSProp "
                                       "(TensorFlow itself is not installable offline)" % (args.steps, batch)},
            "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import topkrec
    from topkrec import dist as tdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload (identical on every rank; each rank samples only the users it owns)
    tr_users, indptr, pos_idx = synth_interactions()
    mine = tdist.user_partition(tr_users, rank, world)
    smp = topkrec.Sampler(mine, indptr, pos_idx, N_ITEMS, seed=123, device=dev)
    st = {k: torch.from_numpy(v).to(dev) for k, v in init_state_np(N_USERS, N_ITEMS, D).items()}
    cfg = topkrec.BprCfg(N_USERS, N_ITEMS, D)
    engine = tdist.DataParallelBpr(cfg, st, B)
    if not args.no_hot_items:   # the items with the most positives get their gradients summed per thread block in shared memory
        topkrec.bpr_set_hot_items(cfg, B, engine.ws, topkrec.popular_items(pos_idx, N_ITEMS))
    POOL = 16                                   # distinct batches cycled through: 16 * 12 B * B = 201 MB > L2
    pool = [topkrec.bpr_sample(smp, (rank * POOL + p) * B, B, dev) for p in range(POOL)]
    loss = torch.zeros(1, dtype=torch.float32, device=dev)

    def step(t):
        u, i, j = pool[t % POOL]
        engine.step(u, i, j, loss=loss)

    for t in range(W):
        step(t)
    barrier()
    topkrec.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for t in range(K):
            step(W + t)
        e1.record()
        barrier()
    launches = topkrec.launch_count()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    value = world * B * K / (ms * K / 1e3)

    # ---- e2e: triples in pinned host memory; every step's triples cross PCIe inside the timed region and every step's
    # loss comes back.  Single GPU: one tkr_bpr_step_host call runs the K steps (the train-loop seam: the copy of step
    # t+1 overlaps the kernels of step t on a side stream); the fully synchronous one-call-per-step figure (the literal
    # sess.run seam) is reported next to it.  Multi GPU: copy + data-parallel step + loss read-back per step.
    loss_h = torch.zeros(max(K, 1), dtype=torch.float32).pin_memory()
    dstage = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3)]
    hpool = [tuple(x.cpu().pin_memory() for x in pool[p]) for p in range(min(POOL, 4))]
    e2e_extra = {}
    if world == 1:
        hk = [torch.cat([pool[(W + t) % POOL][c] for t in range(K)]).cpu().pin_memory() for c in range(3)]
        staging = torch.empty(4 * (K * B * 4 + 256), dtype=torch.uint8, device=dev)

        def run_pipelined():
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], hk[0], hk[1], hk[2], B, K, loss_h,
                                  staging, engine.ws)
        run_pipelined()                                   # warm-up (creates the side stream, touches the pinned pages)
        barrier()
        t0 = time.perf_counter()
        run_pipelined()                                   # synchronises before returning
        e2e_s = time.perf_counter() - t0
        for t in range(W):
            u, i, j = hpool[t % len(hpool)]
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, 1, loss_h, staging, engine.ws)
        t0 = time.perf_counter()
        for t in range(K):
            u, i, j = hpool[t % len(hpool)]
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, 1, loss_h, staging, engine.ws)
        e2e_extra = {"one_call_per_step": B * K / (time.perf_counter() - t0)}
        api = "tkr_bpr_step_host, K steps per call (copies of step t+1 overlap step t); one_call_per_step = the literal sess.run seam of single/bpr.py:141"
    else:
        def e2e_step(t):
            u, i, j = hpool[t % len(hpool)]
            for dst, src in zip(dstage, (u, i, j)):
                dst.copy_(src, non_blocking=True)
            loss.zero_()
            engine.step(*dstage, loss=loss)
            loss_h[:1].copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for t in range(W):
            e2e_step(t)
        barrier()
        t0 = time.perf_counter()
        for t in range(K):
            e2e_step(t)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        api = "host triples -> device, tkr_bpr_grad + all-reduce + tkr_bpr_apply, loss read-back, per step"
    e2e = {"value": world * B * K / e2e_s, "unit": "triples/s", "h2d_bytes_per_step": 12 * B * world,
           "d2h_bytes_per_step": 4 * world, "api": api}
    e2e.update(e2e_extra)

    peaks, peak_src = measured_peaks()
    abytes = algorithmic_bytes_per_triple(D) * B
    achieved = abytes / (ms / 1e3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "triples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BPR synthetic MovieLens-10M shape (70k users x 10k items), d=128, RMSProp, batch %d per GPU" % B,
                       "n_users": N_USERS, "n_items": N_ITEMS, "d": D, "batch_size": B, "optimizer": "rmsprop",
                       "parallelism": "dp%d: users partitioned, V/b replicated, all-reduce of item gradients" % world if world > 1 else "single GPU",
                       "l2": "triple stream cycles through %d pre-sampled batches (%.0f MB) > L2; the %.0f MB of model state "
                             "is reused every step and may stay L2-resident, as in real training" % (POOL, POOL * 12 * B / 1e6, 2 * 4 * D * (N_USERS + N_ITEMS) / 1e6)},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "bpr_grad_kernel + bpr_apply_kernel (one step)", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_launch": abytes, "traffic": profile_traffic("bpr_step")}}

    if world == 1 and not args.skip_sweep:
        line["sweep"] = bpr_sweep(cfg, st, smp, dev)
        line["sweep"].append(bpr_hbm_streaming(dev))
        line["vbpr"] = vbpr_points(smp, dev)
        line["als"] = als_points(dev, cpu=not args.skip_cpu)
    if rank == 0 and world == 1 and not args.skip_cpu:
        # bounded CPU sample: ~10-30 s of the OpenMP port on the same workload
        v1, ms1, threads = cpu_bpr_steps(min(B, 1 << 20), 1, 1)
        n = int(max(2, min(40, 15e3 / ms1)))
        v, _, threads = cpu_bpr_steps(min(B, 1 << 20), n, 0)
        line["cpu_baseline"] = {"value": v, "unit": "triples/s", "cores": threads, "kind": "port",
                                "sample": "%d steps of %d triples, OpenMP C port (oracle/bpr_ref.c) of single/bpr.py:71-101" % (n, min(B, 1 << 20))}
    if not args.skip_score:
        line["score_topk"] = bench_score(args, rank, world, dev, barrier, max_over_ranks, peaks, peak_src)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bpr_sweep(cfg, st, smp, dev):
    """Other operating points of the same kernels on C2 (device-timed, triples resident or fused sampler):
    the reference's own batch size 256 (bpr.py:103), 2^16, and 2^20 with the sampler fused into the step."""
    import torch
    import topkrec
    out = []
    for B, fused, n_steps, reps in ((256, False, 512, 3), (256, True, 512, 3), (1 << 16, False, 16, 5), (1 << 20, True, 1, 10)):
        ws = topkrec.bpr_workspace(cfg, B, dev)
        topkrec.bpr_set_hot_items(cfg, B, ws, topkrec.popular_items(smp.pos_idx, N_ITEMS))
        loss = torch.zeros(n_steps, dtype=torch.float32, device=dev)
        trip = (None, None, None) if fused else topkrec.bpr_sample(smp, 7 << 32, B * n_steps, dev)

        def run(k):
            topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], trip[0], trip[1], trip[2], B, n_steps, ws,
                             loss, sampler=smp if fused else None, first_draw=(9 << 32) + k * B * n_steps)
        for k in range(3):
            run(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            run(3 + k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * n_steps)
        out.append({"batch_size": B, "fused_sampler": fused, "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3),
                    "roofline_frac": algorithmic_bytes_per_triple(D) * B / (ms / 1e3) / 1e9 / measured_peaks()[0]["hbm_gbs"]})
    return out


def bpr_hbm_streaming(dev, n_users=6_000_000, n_items=1_000_000, B=1 << 20, reps=10):
    """The same two kernels on tables far larger than L2: 6 M x 128 user rows (3.1 GB + as much for the
    RMSProp slot and the gradient accumulator) and 1 M item rows, uniform random triples, so every row comes from HBM.
    On C2 the 123 MB of state sits in the 126 MB L2 and the algorithmic roofline fraction exceeds 1; this is the
    HBM-bound operating point of the identical code."""
    import torch
    import topkrec
    g = torch.Generator(device=dev); g.manual_seed(11)
    st = {"U": torch.randn(n_users, D, device=dev, generator=g) * 0.01, "V": torch.randn(n_items, D, device=dev, generator=g) * 0.01,
          "b": torch.zeros(n_items, device=dev)}
    for k in ("U", "V", "b"):
        st["ms" + k] = torch.ones_like(st[k])
    cfg = topkrec.BprCfg(n_users, n_items, D)
    ws = topkrec.bpr_workspace(cfg, B, dev)
    pool = [(torch.randint(0, n_users, (B,), device=dev, generator=g, dtype=torch.int32),
             torch.randint(0, n_items, (B,), device=dev, generator=g, dtype=torch.int32),
             torch.randint(0, n_items, (B,), device=dev, generator=g, dtype=torch.int32)) for _ in range(4)]
    loss = torch.zeros(1, dtype=torch.float32, device=dev)

    def run(t):
        u, i, j = pool[t % 4]
        topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, 1, ws, loss)
    for t in range(3):
        run(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(reps):
        run(3 + t)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"batch_size": B, "fused_sampler": False, "tables": "%d users x %d items, d=%d (%.1f GB of parameters + slots + accumulators, uniform triples)"
           % (n_users, n_items, D, 3 * 4 * D * (n_users + n_items) / 1e9), "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3),
           "roofline_frac": algorithmic_bytes_per_triple(D) * B / (ms / 1e3) / 1e9 / measured_peaks()[0]["hbm_gbs"]}
    del st, ws, pool
    torch.cuda.empty_cache()
    return out


def vbpr_points(smp, dev, d_feat=4096, k=128):
    """BASELINE configs[2]: VBPR with a dense 4096-d feature table resident in HBM (70k users x 10k items, k = 64 + 64),
    fused sampler, device-timed: the reference's batch 256 (vbpr.py:76) and 2^16 / 2^20."""
    import torch
    import topkrec
    g = torch.Generator(device=dev); g.manual_seed(2)
    F = torch.randn(N_ITEMS, d_feat, device=dev, generator=g).abs_()
    F /= F.norm(dim=1, keepdim=True)
    h = k // 2
    cfg = topkrec.VbprCfg(N_USERS, N_ITEMS, k, d_feat)
    st = {"U": torch.randn(N_USERS, k, device=dev, generator=g) * 0.01, "V": torch.zeros(N_ITEMS, k, device=dev),
          "rb": torch.zeros(N_ITEMS, device=dev), "bsum": torch.zeros(N_ITEMS, device=dev),
          "E": torch.full((d_feat, h), 2.0 / (d_feat * k), device=dev), "c": torch.zeros(d_feat, device=dev)}
    st["V"][:, :h] = torch.randn(N_ITEMS, h, device=dev, generator=g) * 0.01
    for n, m in (("U", "msU"), ("V", "msV"), ("rb", "msrb"), ("E", "msE"), ("c", "msc")):
        st[m] = torch.ones_like(st[n])
    out = []
    for B, n_steps, reps in ((256, 64, 3), (1 << 16, 4, 3), (1 << 20, 1, 5)):
        ws = topkrec.vbpr_workspace(cfg, B, dev)
        loss = torch.zeros(n_steps, dtype=torch.float32, device=dev)

        def run(r):
            topkrec.vbpr_step(cfg, st, F, None, None, None, B, n_steps, ws, loss, sampler=smp, first_draw=(11 << 32) + r * B * n_steps)
        for r in range(2):
            run(r)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            run(2 + r)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * n_steps)
        out.append({"batch_size": B, "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3),
                    "config": "VBPR %d users x %d items, k=%d (%d + %d), %d-d dense features resident (%.0f MB)" % (N_USERS, N_ITEMS, k, h, h, d_feat, N_ITEMS * d_feat * 4 / 1e6)})
        del ws
    return out


def als_points(dev, d=256, n_users=480189 // 8, n_items=17770, mean_pos=208, reps=3, cpu_rows=8192, cpu=True):
    """BASELINE configs[3] (CER/WMF, Netflix shape 480 189 users x 17 770 items, d=256): one ALS iteration
    (single/cer.py:36-63 without the content terms = the intended single/wmf.py:67-96) on ONE GPU's share of the
    8-GPU run -- 1/8 of the users, all items -- with Zipf item popularity and ~208 positives per user (100 M / 480 k).
    Device-timed per half-step; FLOP are the reference algorithm's (np.dot(Vi.T, Vi) = 2 n d^2 per row, solve = 2/3 d^3);
    the kernel itself computes the lower triangle only.  Bound: the fp32 FMA pipe (nominal 148 SMs x 128 lanes x 2 x
    1.965 GHz = 74.4 TFLOP/s); cpu_baseline = oracle/als_ref.py user half-step on a bounded sample of rows."""
    import torch
    import topkrec
    rng = np.random.default_rng(0)
    cnt = np.maximum(1, rng.poisson(mean_pos, n_users))
    pop = 1.0 / np.arange(1, n_items + 1); cdf = np.cumsum(pop / pop.sum())
    items = np.searchsorted(cdf, rng.random(int(cnt.sum()))).clip(0, n_items - 1).astype(np.int64)
    users = np.repeat(np.arange(n_users, dtype=np.int64), cnt)
    key = np.sort(users * n_items + items); key = key[np.r_[True, key[1:] != key[:-1]]]      # dedup per user
    users, items = key // n_items, key % n_items
    nnz = int(users.size)
    u_ptr = np.zeros(n_users + 1, np.int64); np.cumsum(np.bincount(users, minlength=n_users), out=u_ptr[1:])
    by_i = np.argsort(items, kind="stable")
    i_ptr = np.zeros(n_items + 1, np.int64); np.cumsum(np.bincount(items, minlength=n_items), out=i_ptr[1:])
    u_idx, i_idx = items.astype(np.int32), users[by_i].astype(np.int32)
    us, its = topkrec.AlsSide(u_ptr, u_idx, 4096, dev), topkrec.AlsSide(i_ptr, i_idx, 4096, dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    U = torch.rand(n_users, d, device=dev, generator=g)            # the reference's start (wmf.py:55-56)
    V = torch.rand(n_items, d, device=dev, generator=g)
    a, b, lu, lv = 1.0, 0.01, 0.01, 0.01
    U0, V0 = U.cpu().numpy(), V.cpu().numpy()

    def iteration():
        XX = topkrec.als_gram(V, its.rated_dev, b, lu)
        l_u = topkrec.als_solve_rows(us, V, U, XX, a, b, 0.0, lu)
        ev[1].record()
        XXv = topkrec.als_gram(U, us.rated_dev, b, 0.0)
        l_i = topkrec.als_solve_rows(its, U, V, XXv, a, b, lv, lv, item_loss=True)
        return l_u, l_i
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tu = ti = 0.0
    topkrec.reset_launch_count()
    for r in range(reps + 1):
        ev[0].record()
        l_u, l_i = iteration()
        ev[2].record(); torch.cuda.synchronize()
        if r:                                                       # first iteration = warm-up
            tu += ev[0].elapsed_time(ev[1]) / reps; ti += ev[1].elapsed_time(ev[2]) / reps
    launches = topkrec.launch_count() // (reps + 1)
    fl_gram = 2.0 * nnz * d * d
    fl_u, fl_i = fl_gram + n_users * (2.0 / 3.0) * d ** 3, fl_gram + n_items * (2.0 / 3.0) * d ** 3
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    out = {"config": "WMF/CER ALS, %d users (1/8 of 480 189: one GPU's share of the 8-GPU run) x %d items, d=%d, %d positives (Zipf items), a=1 b=0.01" % (n_users, n_items, d, nnz),
           "user_step_ms": tu, "item_step_ms": ti, "iteration_ms": tu + ti, "user_rows_per_sec": n_users / (tu / 1e3),
           "item_rows_per_sec": n_items / (ti / 1e3), "positives_per_sec": 2 * nnz / ((tu + ti) / 1e3),
           "loss": float(l_u.sum()) + float(l_i.sum()), "gpu_launches_per_iteration": launches, "dtype": "f32",
           "roofline": {"bound": "fp32 fma pipe", "achieved": (fl_u + fl_i) / ((tu + ti) / 1e3) / 1e12, "peak": peak,
                        "peak_source": "nominal: 148 SMs x 128 FMA lanes x 2 x 1.965 GHz", "unit": "TFLOP/s",
                        "frac": (fl_u + fl_i) / ((tu + ti) / 1e3) / 1e12 / peak,
                        "flop_definition": "reference algorithm: 2*nnz*d^2 + 2/3*d^3 per row, per half-step"}}
    if not cpu:
        return out
    # CPU baseline: the oracle user half-step (np.dot + np.linalg.solve per row, cer.py:39-45) on the first rows
    from oracle import als_ref
    i_rated = np.flatnonzero(np.diff(i_ptr) > 0)
    sub_ptr = u_ptr[:cpu_rows + 1].copy()
    fue = U0[:cpu_rows].copy()
    t0 = time.perf_counter()
    als_ref.user_step(fue, V0, sub_ptr, u_idx, i_rated, a, b, lu)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": cpu_rows / dt, "unit": "user rows/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "%d users of the same workload (incl. one shared Gram of %d item rows), numpy restatement oracle/als_ref.py of single/cer.py:36-46" % (cpu_rows, i_rated.size)}
    return out


def score_sweep(dev, nb, NI, k, peaks, reps=5):
    """BASELINE configs[4]: the same step at d = 64 / 256, and at d = 128 with a rated mask of 64 items per user
    (SURVEY.md 8(d) C5), device-timed with the BF16 item table prepared once."""
    import torch
    import topkrec
    res = []
    for d, rated in ((64, 0), (256, 0), (128, 64)):
        g = torch.Generator(device=dev); g.manual_seed(40 + d)
        V = torch.randn(NI, d, device=dev, generator=g) * 0.1
        U = torch.randn(nb, d, device=dev, generator=g) * 0.1
        rp = ri = None
        if rated:
            ri = torch.sort(torch.randint(0, NI, (nb, rated), device=dev, generator=g, dtype=torch.int32), dim=1).values.reshape(-1).contiguous()
            rp = torch.arange(0, (nb + 1) * rated, rated, device=dev, dtype=torch.int64)
        ws = torch.empty(topkrec.lib().tkr_score_topk_tc_workspace_bytes(nb, NI, d, k, 0), dtype=torch.uint8, device=dev)
        nfb = torch.zeros(1, dtype=torch.int32, device=dev)
        for t in range(3):
            topkrec.score_topk(U, V, k, None, rp, ri, engine="tc", ws=ws, items_prepared=t > 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(reps):
            topkrec.score_topk(U, V, k, None, rp, ri, engine="tc", ws=ws, items_prepared=True, n_fallback=nfb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2.0 * nb * NI * d / (ms / 1e3) / 1e12
        res.append({"d": d, "rated_per_user": rated, "ms_per_step": ms, "users_per_sec": nb / (ms / 1e3), "tflops": tf,
                    "roofline_frac": tf / peaks["bf16_tflops"], "rows_redone_by_exact_fallback": int(nfb.item())})
        del V, U, ws
        torch.cuda.empty_cache()
    return res


def bench_score(args, rank, world, dev, barrier, max_over_ranks, peaks, peak_src):
    """score + top-30: 1M items (sharded over the ranks), d=128, user batches of --score-users."""
    import torch
    import topkrec
    from topkrec import dist as tdist
    # item-sharded runs score larger user batches (each GPU sees every user against 1/world of the items): up to 3 x 18 944 =
    # 56 832 users per step, inside the 8 192..65 536 range of SURVEY 8(d); the fixed-batch number is reported next to it
    NI, k, nb = args.score_items, 30, args.score_users * min(world, 3)
    K, W = max(2, min(args.steps, args.score_steps)), 3
    beg, end = tdist.shard_bounds(NI, world)[rank]
    g = torch.Generator(device=dev); g.manual_seed(4)
    Vfull_rows = end - beg
    V = torch.randn(Vfull_rows, D, device=dev, generator=g) * 0.1
    gu = torch.Generator(device=dev); gu.manual_seed(3)
    Ub = [torch.randn(nb, D, device=dev, generator=gu) * 0.1 for _ in range(4)]

    eng = args.score_engine
    need = topkrec.lib().tkr_score_topk_tc_workspace_bytes(nb, Vfull_rows, D, k, 0) if eng == "tc" else topkrec.lib().tkr_score_topk_workspace_bytes(nb, Vfull_rows, D, k)
    wsb = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)

    nfb = torch.zeros(1, dtype=torch.int32, device=dev)

    state = {"prepared": False}   # the evaluator scores many user batches against one item table: BF16 items converted once

    def step(t):
        prep = state["prepared"] and eng == "tc"
        state["prepared"] = True
        if world == 1:
            return topkrec.score_topk(Ub[t % 4], V, k, col_offset=beg, engine=eng, ws=wsb, n_fallback=nfb if eng == "tc" else None,
                                      items_prepared=prep)
        return tdist.sharded_score_topk(Ub[t % 4], V, k, beg, engine=eng, ws=wsb, items_prepared=prep)
    for t in range(W):
        step(t)
    barrier()
    topkrec.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        e0.record()
        for t in range(K):
            step(t)
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    launches = topkrec.launch_count()
    fixed = None
    if world > 1:                       # the same with the single-GPU batch (strong scaling of one fixed step)
        nb1 = args.score_users
        U1 = [u[:nb1].contiguous() for u in Ub]
        for t in range(W + K):
            if t == W:
                barrier(); e0.record()
            tdist.sharded_score_topk(U1[t % 4], V, k, beg, engine=eng, ws=wsb, items_prepared=eng == "tc" and t > 0)
        e1.record(); barrier()
        ms1 = max_over_ranks(e0.elapsed_time(e1)) / K
        fixed = {"users_per_step": nb1, "ms_per_step": ms1, "users_per_sec": nb1 / (ms1 / 1e3)}
        state["prepared"] = False
    flops = 2.0 * nb * NI * D
    out = {"metric": "scored_users_per_sec_top30", "value": nb / (ms / 1e3), "unit": "users/s", "ms_per_step": ms, "steps": K,
           "config": {"workload": "score + top-30, %d users/step x %d items, d=%d, item-sharded over %d GPU(s); V (%.0f MB/GPU) > L2 per step"
                                  % (nb, NI, D, world, Vfull_rows * D * 4 / 1e6)},
           "dtype": "bf16 tcgen05 filter (fp32 accumulate in TMEM) + exact fp32 fma-chain refine; results bit-identical to the fp32 oracle"
                    if eng == "tc" else "f32 (exact fma-chain scores, CUDA cores)",
           "gpu_launches": launches, "clocks": clk.summary(), "rows_redone_by_exact_fallback_last_step": int(nfb.item()),
           "roofline": {"bound": "tensor", "kernel": "score_filter_kernel" if eng == "tc" else "score_topk_kernel",
                        "achieved": flops / world / (ms / 1e3) / 1e12, "peak": peaks["bf16_tflops"], "peak_source": peak_src, "unit": "TFLOP/s",
                        "frac": flops / world / (ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": profile_traffic("score_topk"),
                        "note": "per GPU: FLOP = 2*nu*(ni/world)*d over the whole step (convert + filter + merge + refine + fallback [+ all-gather]), against the burst bf16 peak"},
           "scaling": "item columns sharded over the GPUs; users per step = %d x min(n_gpus, 3)" % args.score_users}
    if fixed is not None:
        out["fixed_batch"] = fixed
    if world == 1 and not args.skip_sweep and eng == "tc":
        out["sweep"] = score_sweep(dev, nb, NI, k, peaks)
    if world == 1:
        # e2e: the evaluator's flow through the public API (topkrec.score_topk_batches): host V copied once per pass,
        # K host user batches uploaded / K list batches downloaded on side streams while the neighbours compute
        Vh = V.cpu().pin_memory()
        Ke = max(K, 16)       # user batches per pass of the item table (a real evaluation has n_users / 18944 of them)
        Uh = torch.cat([Ub[t % 4] for t in range(Ke)]).cpu().pin_memory()
        Vd = torch.empty_like(V)
        scorer = topkrec.BatchScorer(Vfull_rows, D, k, nb, engine=eng, device=dev)     # streams, double buffers, workspace
        oi = (torch.empty((Ke * nb, k), dtype=torch.int32).pin_memory(), torch.empty((Ke * nb, k), dtype=torch.float32).pin_memory())
        scorer.run(Uh[:2 * nb], V, out=(oi[0][:2 * nb], oi[1][:2 * nb]))                      # warm-up
        barrier()
        t0 = time.perf_counter()
        Vd.copy_(Vh, non_blocking=True)
        scorer.run(Uh, Vd, out=oi)                                                           # synchronises before returning
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": nb * Ke / dt, "unit": "users/s", "h2d_bytes_per_step": nb * D * 4 + Vfull_rows * D * 4 // Ke,
                      "d2h_bytes_per_step": nb * k * 8,
                      "api": "topkrec.BatchScorer (evaluate.py flow): V copied once per pass of %d batches, U batches up / lists down "
                             "on side streams" % Ke}
        if rank == 0 and not args.skip_cpu:
            from oracle import topk_ref
            nsamp = 64
            Un, Vn = Ub[0][:nsamp].cpu().numpy(), V.cpu().numpy()
            t0 = time.perf_counter()
            S = np.dot(Un, Vn.T); np.argsort(S, axis=1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nsamp / dt, "unit": "users/s", "cores": os.cpu_count(), "kind": "reference",
                                   "sample": "np.dot + np.argsort (evaluate.py:78,81 verbatim) on %d users x %d items" % (nsamp, NI)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--score-users", type=int, default=18944)      # 148 tiles of 128 users: one CTA per SM
    ap.add_argument("--score-engine", default="tc", choices=["tc", "exact"])
    ap.add_argument("--score-items", type=int, default=1 << 20)
    ap.add_argument("--score-steps", type=int, default=5)
    ap.add_argument("--no-hot-items", action="store_true", help="do not privatise the most popular item rows (tkr_bpr_workspace_set_hot_items)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-score", action="store_true")
    ap.add_argument("--skip-sweep", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG set in the environment) go to a file
    os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(os.environ.get("TMPDIR", "/tmp"), "nccl_debug.%h.%p.log"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus != world and world == 1 and args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (WORLD_SIZE=%d); running 1 GPU" % (args.gpus, world), file=sys.stderr)
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, local, world)


if __name__ == "__main__":
    main()
