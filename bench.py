#!/usr/bin/env python3
"""Benchmark of the two hot paths on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (BASELINE.json configs[1]): BPR synthetic MovieLens-10M shape, 70k users x 10k items, d=128, reference
optimiser (RMSProp).  One bench "step" = --inner (256) consecutive synchronous mini-batch steps of --batch (2^20)
triples per GPU, so that the K timed steps last >= 2 s whatever K the driver passes.  `value` =
triples/s with the triples resident in HBM; `e2e` = the same through tkr_bpr_step_host (triples in pinned host memory,
H2D + loss D2H inside the timed region).  `roofline` is the HBM-streaming operating point of the same kernels (tables
far beyond L2), measured live in the same run: on C2 the 123 MB of state is L2-resident and an HBM fraction means
nothing there (reported as roofline.c2 against a measured L2 copy rate).  The second path (score + top-30, BASELINE
configs[4] slice: 18 944 users x 1 M items, d=128) is reported in the same line under "score_topk"; VBPR (configs[2])
and ALS (configs[3]) points under "vbpr" / "als".  `--impl reference` times the CPU port of the reference step
(oracle/bpr_ref.c, OpenMP, all host threads) on the same workload.
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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
    """Samples SM clock + throttle reasons during a timed region (pynvml, 5 ms period)."""
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
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel):
    """DRAM bytes per launch from the committed ncu captures (profiles/traffic.json names the capture), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


def device_time_ms(fn, reps, warm=2):
    """fn() reps times between two CUDA events on the current stream (after `warm` untimed calls and a synchronise)"""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measured_fp32_tflops(dev):
    """cuBLAS SGEMM 8192^3 with TF32 off: the FMA-pipe roof used for the fp32 kernels (ALS, VBPR content GEMMs)"""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
        ms = min(device_time_ms(lambda: torch.matmul(a, b), 3, 1) for _ in range(3))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return 2 * 8192 ** 3 / (ms / 1e3) / 1e12


# ----------------------------------------------------------------------------- CPU port (reference arm / cpu_baseline)
def cpu_bpr_steps(batch, n_steps, warmup, seed=123):
    """Time the OpenMP C port of the reference step (oracle/bpr_ref.c) on the C2 workload with every host thread this
    process may use (torchrun exports OMP_NUM_THREADS=1: the team size is set explicitly).
    Returns (triples_per_sec, ms_per_step, threads)."""
    from oracle import bpr_ref, clib
    threads = int(clib.lib().tkr_ref_omp_threads(host_threads()))
    st = init_state_np(N_USERS, N_ITEMS, D)
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, N_ITEMS + 1); pop /= pop.sum()
    u = rng.integers(0, N_USERS, batch).astype(np.int32)
    i = rng.choice(N_ITEMS, batch, p=pop).astype(np.int32)
    j = rng.integers(0, N_ITEMS, batch).astype(np.int32)
    cfg = bpr_ref.BprCfg()
    for _ in range(warmup):
        bpr_ref.c_bpr_train(st, u, i, j, batch, cfg)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        bpr_ref.c_bpr_train(st, u, i, j, batch, cfg)
    dt = time.perf_counter() - t0
    return batch * n_steps / dt, 1e3 * dt / n_steps, threads


def workload_name(batch, inner, world=1):
    return ("BPR synthetic MovieLens-10M shape (70k users x 10k items), d=128, RMSProp; one step = %d synchronous mini-batches of %d triples%s"
            % (inner, batch, " per GPU" if world > 1 else ""))


def run_reference(args, rank):
    if rank != 0:
        return
    batch = min(args.batch, 1 << 20)
    val, ms, threads = cpu_bpr_steps(batch, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "triples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(batch, args.inner), "n_users": N_USERS, "n_items": N_ITEMS, "d": D, "batch_size": batch,
                       "sample": "each timed step = ONE mini-batch of %d triples (1/%d of the GPU arm's step), same tables, same optimiser" % (batch, args.inner)},
            "cpu_baseline": {"value": val, "unit": "triples/s", "cores": threads, "kind": "port",
                             "sample": "%d steps of %d triples; OpenMP C port of single/bpr.py:71-101 + TF-1.15 RMSProp on %d threads "
                                       "(TensorFlow itself is not installable offline)" % (args.steps, batch, threads)},
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
    B, K, W, INNER = args.batch, args.steps, args.warmup, args.inner

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
    engine = tdist.DataParallelBpr(cfg, st, B, exchange=args.exchange if world > 1 else None)
    if not args.no_hot_items:   # the items with the most positives get their gradients summed per thread block in shared memory
        topkrec.bpr_set_hot_items(cfg, B, engine.ws, topkrec.popular_items(pos_idx, N_ITEMS))
    POOL = 16                                   # distinct batches cycled through: 16 * 12 B * B = 201 MB > L2
    pool = [topkrec.bpr_sample(smp, (rank * POOL + p) * B, B, dev) for p in range(POOL)]
    loss = torch.zeros(1, dtype=torch.float32, device=dev)

    def step(t):                                # one bench step = INNER consecutive synchronous mini-batch steps
        for m in range(INNER):
            u, i, j = pool[(t * INNER + m) % POOL]
            engine.step(u, i, j, loss=loss)

    for t in range(W):
        step(t)
    engine.check()
    barrier()
    topkrec.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for t in range(K):
            step(W + t)
        e1.record()
        barrier()
    engine.check()
    launches = topkrec.launch_count()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    value = world * B * INNER / (ms / 1e3)
    dp_check = dp_consistency(engine, st, dev, world) if world > 1 else None

    # ---- e2e: triples in pinned host memory; every mini-batch's triples cross PCIe inside the timed region and every
    # mini-batch's loss comes back.  Single GPU: one tkr_bpr_step_host call per bench step (the train-loop seam: the
    # copy of mini-batch t+1 overlaps the kernels of mini-batch t on a side stream); the fully synchronous
    # one-call-per-mini-batch figure (the literal sess.run seam) is reported next to it.  Multi GPU: the same pipeline
    # driven from Python around the fused data-parallel step.
    EI = min(INNER, args.e2e_inner)
    e2e_extra = {}
    hk = [torch.cat([pool[m % POOL][c] for m in range(EI)]).cpu().pin_memory() for c in range(3)]
    loss_h = torch.zeros(EI, dtype=torch.float32).pin_memory()
    if world == 1:
        staging = torch.empty(4 * (EI * B * 4 + 256), dtype=torch.uint8, device=dev)

        def e2e_step():
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], hk[0], hk[1], hk[2], B, EI, loss_h,
                                  staging, engine.ws)
        for _ in range(max(1, min(W, 2))):
            e2e_step()                                    # warm-up (creates the side stream, touches the pinned pages)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            e2e_step()                                    # synchronises before returning
        e2e_s = time.perf_counter() - t0
        hp = [tuple(x[:B] for x in hk)]
        for _ in range(3):
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], *hp[0], B, 1, loss_h, staging, engine.ws)
        n1 = max(20, min(200, K * 4))
        t0 = time.perf_counter()
        for _ in range(n1):
            topkrec.bpr_step_host(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], *hp[0], B, 1, loss_h, staging, engine.ws)
        e2e_extra = {"one_call_per_minibatch": B * n1 / (time.perf_counter() - t0)}
        api = ("tkr_bpr_step_host: one call per bench step, %d mini-batches per call (the copy of mini-batch t+1 overlaps mini-batch t); "
               "one_call_per_minibatch = the literal sess.run seam of single/bpr.py:141" % EI)
    else:
        copy = torch.cuda.Stream(dev)
        dbuf = [[torch.empty(B, dtype=torch.int32, device=dev) for _ in range(3)] for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        lbuf = torch.zeros(EI, dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream(dev)

        def e2e_step():
            lbuf.zero_()
            for m in range(EI):
                s = m & 1
                with torch.cuda.stream(copy):
                    if m >= 2:
                        copy.wait_event(free[s])
                    for c in range(3):
                        dbuf[s][c].copy_(hk[c][m * B:(m + 1) * B], non_blocking=True)
                    ready[s].record(copy)
                main.wait_event(ready[s])
                engine.step(*dbuf[s], loss=lbuf[m:m + 1])
                free[s].record(main)
            loss_h.copy_(lbuf, non_blocking=True)
            main.synchronize()
        copy.wait_stream(main)
        for _ in range(max(1, min(W, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        engine.check()
        api = "pinned host triples -> device on a copy stream (double buffered), tkr_bpr_dp_step per mini-batch, losses read back per bench step (%d mini-batches)" % EI
    e2e = {"value": world * B * EI * K / e2e_s, "unit": "triples/s", "h2d_bytes_per_step": 12 * B * EI * world,
           "d2h_bytes_per_step": 4 * EI * world, "minibatches_per_step": EI, "api": api}
    e2e.update(e2e_extra)

    peaks, peak_src = measured_peaks()
    abytes = algorithmic_bytes_per_triple(D) * B
    ms_mb = ms / INNER                                              # one mini-batch = one launch pair
    achieved_c2 = abytes / (ms_mb / 1e3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "triples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(B, INNER, world), "n_users": N_USERS, "n_items": N_ITEMS, "d": D, "batch_size": B,
                       "minibatches_per_step": INNER, "ms_per_minibatch": ms_mb, "optimizer": "rmsprop",
                       "parallelism": ("dp%d: users partitioned, V/b replicated; item gradients exchanged and applied by ONE kernel over peer memory "
                                       "(NVLink loads/stores, tkr_bpr_dp_step)" % world if engine.exchange == "peer" else
                                       "dp%d: users partitioned, V/b replicated, NCCL all-reduce of item gradients" % world) if world > 1 else "single GPU",
                       "l2": "triple stream cycles through %d pre-sampled batches (%.0f MB) > L2; the %.0f MB of model state "
                             "is reused every step and stays L2-resident, as in real training at this shape" % (POOL, POOL * 12 * B / 1e6, 2 * 4 * D * (N_USERS + N_ITEMS) / 1e6)},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches}
    if dp_check is not None:
        line["multi_gpu_check"] = dp_check

    # ---- roofline: the HBM-bound operating point of the same two kernels, measured live on rank 0
    if rank == 0:
        stream_pt = bpr_hbm_streaming(dev)
        line["roofline"] = {"bound": "hbm", "kernel": "bpr_count_kernel + bpr_grad_kernel<INPLACE> + bpr_apply_kernel (one mini-batch step)",
                            "workload": stream_pt["tables"], "achieved": stream_pt["algorithmic_gbs"], "peak": peaks["hbm_gbs"], "peak_source": peak_src,
                            "unit": "GB/s", "frac": stream_pt["algorithmic_gbs"] / peaks["hbm_gbs"], "ms_per_launch": stream_pt["us_per_step"] / 1e3,
                            "algorithmic_bytes_per_launch": abytes, "traffic": profile_traffic("bpr_step_streaming"),
                            "traffic_source": "ncu --set full of this configuration, profiles/r02a_ncu_bpr_streaming.txt (dram read + write of the three launches of a step)",
                            "note": "tables far beyond L2: every row comes from HBM.  The headline workload (C2, 123 MB of state) is L2-resident; its rate is in roofline.c2",
                            "c2": {"achieved_algorithmic_gbs": achieved_c2, "frac_of_hbm_peak": achieved_c2 / peaks["hbm_gbs"],
                                   "l2_resident": True, "dram_traffic_per_minibatch": profile_traffic("bpr_step"),
                                   "note": "algorithmic bytes (duplicates counted per triple) / mini-batch time; > HBM peak because parameters, slots and accumulators "
                                           "stay in the 126 MB L2 and the ~200 occurrences of a popular item row per mini-batch are served on chip (ncu: 0.32 GB of DRAM traffic per 6.49 GB "
                                           "algorithmic); what binds there is gather latency and L2 atomic throughput (profiles/r01p_ncu_bpr.txt), not a bandwidth roof"}}
    barrier()

    if world == 1 and not args.skip_sweep:
        line["sweep"] = bpr_sweep(cfg, st, smp, dev)
        line["sweep"].append(bpr_hbm_streaming(dev, n_users=16_000_000, n_items=8_000_000))
        fp32_peak = measured_fp32_tflops(dev)
        line["vbpr"] = vbpr_points(smp, dev, fp32_peak)
        line["als"] = als_points(dev, fp32_peak, cpu=not args.skip_cpu)
    if world > 1 and not args.skip_sweep:
        line["vbpr"] = vbpr_dp_point(smp, dev, rank, world, barrier, max_over_ranks)
        line["als"] = als_sharded_point(dev, rank, world, barrier, max_over_ranks)
    if rank == 0 and world == 1 and not args.skip_cpu:
        # bounded CPU sample: ~10-30 s of the OpenMP port on the same workload
        v1, ms1, threads = cpu_bpr_steps(min(B, 1 << 20), 1, 1)
        n = int(max(2, min(40, 15e3 / ms1)))
        v, _, threads = cpu_bpr_steps(min(B, 1 << 20), n, 0)
        line["cpu_baseline"] = {"value": v, "unit": "triples/s", "cores": threads, "kind": "port",
                                "sample": "%d mini-batch steps of %d triples, OpenMP C port (oracle/bpr_ref.c) of single/bpr.py:71-101 on %d threads" % (n, min(B, 1 << 20), threads)}
    if not args.skip_score:
        line["score_topk"] = bench_score(args, rank, world, dev, barrier, max_over_ranks, peaks, peak_src)
    engine.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dp_consistency(engine, st, dev, world):
    """N > 1: the replicas of V / b must hold identical bits after the timed steps (the fused exchange writes every
    replica from the row's owner)."""
    import torch
    import torch.distributed as dist
    out = {}
    for k in ("V", "b"):
        hi, lo = st[k].clone(), st[k].clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        out["replicas_bit_identical_" + k] = bool(torch.equal(hi, lo))
    out["finite"] = bool(torch.isfinite(st["V"]).all().item() and torch.isfinite(st["U"]).all().item())
    out["equals_one_gpu_union_batch"] = "tests + profiles/dp_ngpu.py (<= 1e-6 relative at 2 and 8 GPUs)"
    return out


def bpr_sweep(cfg, st, smp, dev):
    """Other operating points of the same path on C2 (device-timed, >= 1 s each): the reference's own batch size 256
    (bpr.py:103; persistent multi-step cluster kernel), 64 and 1024 (dataflow multi-step kernel), 2^16, and 2^20 with the
    sampler fused into the step."""
    import torch
    import topkrec
    out = []
    for B, fused, n_steps in ((64, False, 4096), (256, False, 4096), (256, True, 4096), (1024, False, 1024), (1 << 16, False, 64), (1 << 20, True, 8)):
        ws = topkrec.bpr_workspace(cfg, B, dev)
        topkrec.bpr_set_hot_items(cfg, B, ws, topkrec.popular_items(smp.pos_idx, N_ITEMS))
        loss = torch.zeros(n_steps, dtype=torch.float32, device=dev)
        trip = (None, None, None) if fused else topkrec.bpr_sample(smp, 7 << 32, B * n_steps, dev)
        state = {"k": 0}

        def run():
            state["k"] += 1
            topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], trip[0], trip[1], trip[2], B, n_steps, ws,
                             loss, sampler=smp if fused else None, first_draw=(9 << 32) + state["k"] * B * n_steps)
        one = device_time_ms(run, 2, 2)
        reps = int(max(3, min(400, 1000.0 / one)))
        ms = device_time_ms(run, reps, 0) / n_steps
        out.append({"batch_size": B, "fused_sampler": fused, "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3), "timed_seconds": ms * n_steps * reps / 1e3,
                    "route": ("dataflow multi-step kernel (row-level version words, no grid barriers)" if B <= 64 or 256 < B <= 1024 else
                              "persistent cluster kernel, %d steps per launch" % n_steps if B <= 256 else "bpr_grad_kernel + bpr_apply_kernel per step"),
                    "algorithmic_gbs": algorithmic_bytes_per_triple(D) * B / (ms / 1e3) / 1e9})
    return out


def bpr_hbm_streaming(dev, n_users=6_000_000, n_items=1_000_000, B=1 << 20, seconds=1.0):
    """The same kernels on tables far larger than L2: uniform random triples over n_users x 128 user rows and
    n_items item rows (each with its RMSProp slot row and gradient-accumulator row), so every row comes from HBM.  On C2 the
    123 MB of state sits in the 126 MB L2; this is the HBM-bound operating point of the identical code path."""
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
    state = {"t": 0}

    def run():
        u, i, j = pool[state["t"] % 4]
        state["t"] += 1
        topkrec.bpr_step(cfg, st["U"], st["V"], st["b"], st["msU"], st["msV"], st["msb"], u, i, j, B, 1, ws, loss)
    one = device_time_ms(run, 3, 3)
    reps = int(max(5, min(2000, seconds * 1e3 / one)))
    ms = device_time_ms(run, reps, 0)
    gbs = algorithmic_bytes_per_triple(D) * B / (ms / 1e3) / 1e9
    out = {"batch_size": B, "fused_sampler": False, "tables": "%d users x %d items, d=%d (%.1f GB of parameters + slots + accumulators), uniform triples, B=%d"
           % (n_users, n_items, D, 3 * 4 * D * (n_users + n_items) / 1e9, B), "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3),
           "timed_seconds": ms * reps / 1e3, "algorithmic_gbs": gbs, "roofline_frac": gbs / measured_peaks()[0]["hbm_gbs"]}
    del st, ws, pool
    torch.cuda.empty_cache()
    return out


def vbpr_points(smp, dev, fp32_peak, d_feat=4096, k=128):
    """BASELINE configs[2]: VBPR with a dense 4096-d feature table resident in HBM (70k users x 10k items, k = 64 + 64),
    fused sampler, device-timed: the reference's batch 256 (vbpr.py:76) and 2^16 / 2^20.  The content part of a step is two
    GEMMs over all touched items, F.[E|c] (projection) and F^T.[W|wq] (dE, dc): 2 * 2 * n_items * d_feat * (k/2 + 1) FLOP."""
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
    for B, n_steps in ((256, 256), (1 << 16, 16), (1 << 20, 4)):
        ws = topkrec.vbpr_workspace(cfg, B, dev)
        topkrec.vbpr_set_hot_items(cfg, B, ws, topkrec.popular_items(smp.pos_idx, N_ITEMS))
        loss = torch.zeros(n_steps, dtype=torch.float32, device=dev)
        state = {"r": 0}

        def run():
            state["r"] += 1
            topkrec.vbpr_step(cfg, st, F, None, None, None, B, n_steps, ws, loss, sampler=smp, first_draw=(11 << 32) + state["r"] * B * n_steps)
        one = device_time_ms(run, 2, 2)
        reps = int(max(3, min(200, 1000.0 / one)))
        ms = device_time_ms(run, reps, 0) / n_steps
        touched = min(2 * B, N_ITEMS)
        gemm_flop = 2.0 * 2.0 * touched * d_feat * (h + 1)
        out.append({"batch_size": B, "us_per_step": 1e3 * ms, "triples_per_sec": B / (ms / 1e3), "timed_seconds": ms * n_steps * reps / 1e3,
                    "config": "VBPR %d users x %d items, k=%d (%d + %d), %d-d dense features resident (%.0f MB)" % (N_USERS, N_ITEMS, k, h, h, d_feat, N_ITEMS * d_feat * 4 / 1e6),
                    "roofline": {"bound": "fp32 fma pipe (content GEMMs) + hbm (gather/scatter part)", "content_gemm_flop_per_step": gemm_flop,
                                 "achieved_tflops_whole_step": gemm_flop / (ms / 1e3) / 1e12, "peak": fp32_peak,
                                 "peak_source": "measured live: cuBLAS SGEMM 8192^3, TF32 off", "unit": "TFLOP/s",
                                 "frac": gemm_flop / (ms / 1e3) / 1e12 / fp32_peak,
                                 "note": "whole step time charged to the content GEMM FLOP (<= 2B item rows projected / differentiated per step)"}})
        del ws
    return out


def vbpr_dp_point(smp, dev, rank, world, barrier, max_over_ranks, d_feat=4096, k=128, B=1 << 20):
    """BASELINE configs[2] at N GPUs: data-parallel VBPR (topkrec.dist.DataParallelVbpr: users partitioned, item / content tables
    replicated, all-reduce of the item-side gradients and of dE / dc), 2^20 triples per GPU per step (weak scaling)."""
    import torch
    import torch.distributed as dist
    import topkrec
    from topkrec import dist as tdist
    g = torch.Generator(device=dev); g.manual_seed(2)              # the same tables on every rank
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
    eng = tdist.DataParallelVbpr(cfg, st, F, B)
    topkrec.vbpr_set_hot_items(cfg, B, eng.ws, topkrec.popular_items(smp.pos_idx, N_ITEMS))
    topkrec.vbpr_project(cfg, st, F)
    state = {"r": 0}

    def run():
        state["r"] += 1
        eng.step(sampler=smp, first_draw=(13 << 32) + state["r"] * B)
    for _ in range(5):
        run()
    barrier()
    reps = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    hi, lo = st["E"].clone(), st["E"].clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return [{"batch_size": B, "n_gpus": world, "us_per_step": 1e3 * ms, "triples_per_sec": world * B / (ms / 1e3), "timed_seconds": ms * reps / 1e3, "scaling": "weak",
             "replicas_bit_identical_E": bool(torch.equal(hi, lo)),
             "config": "VBPR %d users x %d items, k=%d, %d-d dense features, data-parallel over %d GPUs: NCCL all-reduce of [GV|Gb|tchV] (%.1f MB) and [GE|Gc] (%.1f MB) per step"
                       % (N_USERS, N_ITEMS, k, d_feat, world, eng.sparse.numel() * 4 / 1e6, eng.dense.numel() * 4 / 1e6)}]


def als_workload(n_users, n_items, mean_pos):
    rng = np.random.default_rng(0)
    cnt = np.maximum(1, rng.poisson(mean_pos, n_users))
    pop = 1.0 / np.arange(1, n_items + 1); cdf = np.cumsum(pop / pop.sum())
    items = np.searchsorted(cdf, rng.random(int(cnt.sum()))).clip(0, n_items - 1).astype(np.int64)
    users = np.repeat(np.arange(n_users, dtype=np.int64), cnt)
    key = np.sort(users * n_items + items); key = key[np.r_[True, key[1:] != key[:-1]]]      # dedup per user
    users, items = key // n_items, key % n_items
    u_ptr = np.zeros(n_users + 1, np.int64); np.cumsum(np.bincount(users, minlength=n_users), out=u_ptr[1:])
    by_i = np.argsort(items, kind="stable")
    i_ptr = np.zeros(n_items + 1, np.int64); np.cumsum(np.bincount(items, minlength=n_items), out=i_ptr[1:])
    return u_ptr, items.astype(np.int32), i_ptr, users[by_i].astype(np.int32), int(users.size)


def als_points(dev, fp32_peak, d=256, n_users=480189 // 8, n_items=17770, mean_pos=208, reps=3, cpu_rows=8192, cpu=True):
    """BASELINE configs[3] (CER/WMF, Netflix shape 480 189 users x 17 770 items, d=256): one ALS iteration
    (single/cer.py:36-63 without the content terms = the intended single/wmf.py:67-96) on ONE GPU's share of the
    8-GPU run -- 1/8 of the users, all items -- with Zipf item popularity and ~208 positives per user (100 M / 480 k).
    Device-timed per half-step; FLOP are the reference algorithm's (np.dot(Vi.T, Vi) = 2 n d^2 per row, solve = 2/3 d^3);
    the kernel itself computes the lower triangle only.  Bound: the fp32 FMA pipe, against a live cuBLAS SGEMM rate;
    cpu_baseline = oracle/als_ref.py user half-step on a bounded sample of rows."""
    import torch
    import topkrec
    u_ptr, u_idx, i_ptr, i_idx, nnz = als_workload(n_users, n_items, mean_pos)
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
    with ClockSampler(dev.index) as clk:                            # (this section follows >= 1 s of tensor-pipe work: the clocks say in what state)
        for r in range(reps + 1):
            ev[0].record()
            l_u, l_i = iteration()
            ev[2].record(); torch.cuda.synchronize()
            if r:                                                   # first iteration = warm-up
                tu += ev[0].elapsed_time(ev[1]) / reps; ti += ev[1].elapsed_time(ev[2]) / reps
    launches = topkrec.launch_count() // (reps + 1)
    fl_gram = 2.0 * nnz * d * d
    fl_u, fl_i = fl_gram + n_users * (2.0 / 3.0) * d ** 3, fl_gram + n_items * (2.0 / 3.0) * d ** 3
    out = {"config": "WMF/CER ALS, %d users (1/8 of 480 189: one GPU's share of the 8-GPU run) x %d items, d=%d, %d positives (Zipf items), a=1 b=0.01" % (n_users, n_items, d, nnz),
           "user_step_ms": tu, "item_step_ms": ti, "iteration_ms": tu + ti, "user_rows_per_sec": n_users / (tu / 1e3),
           "item_rows_per_sec": n_items / (ti / 1e3), "positives_per_sec": 2 * nnz / ((tu + ti) / 1e3),
           "loss": float(l_u.sum()) + float(l_i.sum()), "gpu_launches_per_iteration": launches, "dtype": "f32", "clocks": clk.summary(),
           "roofline": {"bound": "fp32 fma pipe", "achieved": (fl_u + fl_i) / ((tu + ti) / 1e3) / 1e12, "peak": fp32_peak,
                        "peak_source": "measured live: cuBLAS SGEMM 8192^3, TF32 off", "unit": "TFLOP/s",
                        "frac": (fl_u + fl_i) / ((tu + ti) / 1e3) / 1e12 / fp32_peak,
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
    out["cpu_baseline"] = {"value": cpu_rows / dt, "unit": "user rows/s", "cores": host_threads(), "kind": "port",
                           "sample": "%d users of the same workload (incl. one shared Gram of %d item rows), numpy restatement oracle/als_ref.py of single/cer.py:36-46" % (cpu_rows, i_rated.size)}
    return out


def als_sharded_point(dev, rank, world, barrier, max_over_ranks, d=256, n_items=17770, mean_pos=208, reps=3):
    """BASELINE configs[3] at N GPUs: WMF/CER ALS iteration with rows sharded over the ranks (topkrec.dist.ShardedAls:
    every rank solves its block of users, then of items; solved blocks are exchanged; replicas stay bit-identical).
    Weak in the users (60 023 per GPU: 480 189 at 8 GPUs), all 17 770 items."""
    import torch
    import torch.distributed as dist
    from topkrec import dist as tdist
    n_users = (480189 // 8) * world
    u_ptr, u_idx, i_ptr, i_idx, nnz = als_workload(n_users, n_items, mean_pos)
    eng = tdist.ShardedAls(u_ptr, u_idx, i_ptr, i_idx, seg=4096, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    U = torch.rand(n_users, d, device=dev, generator=g)
    V = torch.rand(n_items, d, device=dev, generator=g)
    a, b, lu, lv = 1.0, 0.01, 0.01, 0.01
    eng.iteration(U, V, a, b, lu, lv, wmf=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        losses = eng.iteration(U, V, a, b, lu, lv, wmf=True)
    e1.record(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    hi, lo = U.clone(), U.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return {"config": "WMF/CER ALS, %d users x %d items, d=%d, %d positives, rows sharded over %d GPUs (weak in the users)" % (n_users, n_items, d, nnz, world),
            "iteration_ms": ms, "user_rows_per_sec": n_users / (ms / 1e3), "positives_per_sec": 2 * nnz / (ms / 1e3), "loss": float(losses[0] + losses[1]),
            "replicas_bit_identical": bool(torch.equal(hi, lo)), "scaling": "weak"}


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
        topkrec.score_topk(U, V, k, None, rp, ri, engine="tc", ws=ws, items_prepared=False)
        one = device_time_ms(lambda: topkrec.score_topk(U, V, k, None, rp, ri, engine="tc", ws=ws, items_prepared=True, n_fallback=nfb), 3, 2)
        reps = int(max(5, min(400, 1000.0 / one)))
        ms = device_time_ms(lambda: topkrec.score_topk(U, V, k, None, rp, ri, engine="tc", ws=ws, items_prepared=True, n_fallback=nfb), reps, 0)
        tf = 2.0 * nb * NI * d / (ms / 1e3) / 1e12
        res.append({"d": d, "rated_per_user": rated, "ms_per_step": ms, "users_per_sec": nb / (ms / 1e3), "tflops": tf, "timed_seconds": ms * reps / 1e3,
                    "roofline_frac": tf / peaks["bf16_tflops"], "rows_redone_by_exact_fallback": int(nfb.item())})
        del V, U, ws
        torch.cuda.empty_cache()
    return res


def bench_score(args, rank, world, dev, barrier, max_over_ranks, peaks, peak_src):
    """score + top-30: 1M items (sharded over the ranks), d=128, user batches of --score-users (the SAME batch at every N:
    strong scaling of one step).  N > 1: topkrec.dist.RingScorer -- one sweep per batch cut into a segment per GPU, the per-row
    filter state travelling rank to rank over NVLink, G batches in flight; the owner of every batch checks its lists and score
    bits against the unsharded engine.  --score-dist exchange: topkrec.dist.ShardedScorer (independent shard lists + peer-memory
    exchange by user slice), kept as the route the ring is measured against."""
    import torch
    import topkrec
    from topkrec import dist as tdist
    NI, k, nb = args.score_items, 30, args.score_users
    beg, end = tdist.shard_bounds(NI, world)[rank]
    g = torch.Generator(device=dev); g.manual_seed(4)
    Vfull = torch.randn(NI, D, device=dev, generator=g) * 0.1      # the same table on every rank; each rank keeps its shard
    V = Vfull[beg:end].contiguous() if world > 1 else Vfull
    gu = torch.Generator(device=dev); gu.manual_seed(3)
    Ub = [torch.randn(nb, D, device=dev, generator=gu) * 0.1 for _ in range(4)]
    eng = args.score_engine
    check = None
    sc = None
    how = "one GPU"
    if world > 1 and args.score_dist == "ring" and eng == "tc":
        # ring of sweep segments: rank r sweeps item shard r, the rows' thresholds / candidate buffers travel rank to rank over
        # NVLink, the rank holding a batch's last segment re-scores exactly against the full table and owns the lists
        how = "ring of sweep segments (topkrec.dist.RingScorer): per-row filter state copied rank to rank over NVLink, G batches in flight"
        sc = tdist.RingScorer(V, D, k, nb, beg, Vfull, device=dev)
        T0 = 2 * world + 1
        bad = []

        def check_owner(t, idx, score):
            wi, wsc = topkrec.score_topk(Ub[t % 4], Vfull, k, engine="tc")
            if not (torch.equal(idx, wi) and torch.equal(score.view(torch.int32), wsc.view(torch.int32))):
                bad.append(t)
        sc.run([Ub[t % 4] for t in range(T0)], on_result=check_owner)
        sc.wait()
        ok = torch.tensor([int(not bad)], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        check = {"ring_lists_and_score_bits_equal_unsharded_on_every_owner": bool(ok.item()), "batches_checked": T0}
        torch.cuda.empty_cache()

        def run_steps(n):
            sc.run([Ub[t % 4] for t in range(n)])
            sc.wait()
    elif world > 1:
        how = "independent per-shard lists, candidates exchanged by user slice over peer memory (topkrec.dist.ShardedScorer)"
        sc = tdist.ShardedScorer(end - beg, D, k, nb, beg, engine=eng, device=dev)
        oi, osc = sc.submit(Ub[0], V)
        sc.wait()
        b0, b1 = sc.rows_of(nb)
        wi, wsc = topkrec.score_topk(Ub[0][b0:b1].contiguous(), Vfull, k, engine=eng)
        ok = torch.tensor([int(torch.equal(oi, wi) and torch.equal(osc.view(torch.int32), wsc.view(torch.int32)))], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        check = {"sharded_lists_and_score_bits_equal_unsharded_on_every_rank": bool(ok.item())}
        del Vfull, wi, wsc
        torch.cuda.empty_cache()

        def run_steps(n):
            for t in range(n):
                sc.submit(Ub[t % 4], V)
            sc.wait()
    else:
        need = topkrec.lib().tkr_score_topk_tc_workspace_bytes(nb, NI, D, k, 0) if eng == "tc" else topkrec.lib().tkr_score_topk_workspace_bytes(nb, NI, D, k)
        wsb = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        nfb = torch.zeros(1, dtype=torch.int32, device=dev)
        state = {"prepared": False}   # the evaluator scores many user batches against one item table: BF16 items converted once

        def run_steps(n):
            for t in range(n):
                prep = state["prepared"] and eng == "tc"
                state["prepared"] = True
                topkrec.score_topk(Ub[t % 4], V, k, col_offset=beg, engine=eng, ws=wsb, n_fallback=nfb if eng == "tc" else None, items_prepared=prep)
    run_steps(2 * world + 2)
    barrier()
    # >= 1 s of steps whatever --steps says (a step is ~4 ms / N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = 8 * world
    e0.record()
    run_steps(n0)
    e1.record(); barrier()
    K = int(max(args.score_steps, min(4000, 1000.0 / (max_over_ranks(e0.elapsed_time(e1)) / n0))))
    K = int(max_over_ranks(float(K)))
    topkrec.reset_launch_count()
    with ClockSampler(dev.index) as clk:
        e0.record()
        run_steps(K)
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / K
    launches = topkrec.launch_count()
    flops = 2.0 * nb * NI * D
    out = {"metric": "scored_users_per_sec_top30", "value": nb / (ms / 1e3), "unit": "users/s", "ms_per_step": ms, "steps": K,
           "config": {"workload": "score + top-30, %d users/step x %d items, d=%d, item-sharded over %d GPU(s); V (%.0f MB/GPU) > L2 per step"
                                  % (nb, NI, D, world, (end - beg) * D * 4 / 1e6)},
           "dtype": "bf16 tcgen05 filter (fp32 accumulate in TMEM) + exact fp32 fma-chain refine; results bit-identical to the fp32 oracle"
                    if eng == "tc" else "f32 (exact fma-chain scores, CUDA cores)",
           "gpu_launches": launches, "clocks": clk.summary(),
           "roofline": {"bound": "tensor", "kernel": "score_filter_kernel" if eng == "tc" else "score_topk_kernel",
                        "achieved": flops / world / (ms / 1e3) / 1e12, "peak": peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                        "peak_source": peak_src + ": sustained cuBLAS bf16 rate (this timed region is >= 1 s of back-to-back tensor work and runs into the power cap, see clocks)",
                        "unit": "TFLOP/s", "frac": flops / world / (ms / 1e3) / 1e12 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                        "frac_of_burst_peak": flops / world / (ms / 1e3) / 1e12 / peaks["bf16_tflops"], "traffic": profile_traffic("score_topk"),
                        "note": "per GPU: FLOP = 2*nu*(ni/world)*d over the whole step (convert + filter + merge + refine + fallback [+ exchange])"},
           "scaling": "strong: the same %d-user batch at every N; item columns sharded over the GPUs -- %s" % (nb, how)}
    if world == 1 and eng == "tc":
        out["rows_redone_by_exact_fallback_last_step"] = int(nfb.item())
    if check is not None:
        out["multi_gpu_check"] = check
        sc.close()
    if world == 1 and not args.skip_sweep and eng == "tc":
        out["sweep"] = score_sweep(dev, nb, NI, k, peaks)
    if world == 1:
        # e2e: the evaluator's flow through the public API (topkrec.BatchScorer): host V copied once per pass,
        # K host user batches uploaded / K list batches downloaded on side streams while the neighbours compute
        Vh = V.cpu().pin_memory()
        Ke = 32               # user batches per pass of the item table (a real evaluation has n_users / 18944 of them)
        Uh = torch.cat([Ub[t % 4] for t in range(Ke)]).cpu().pin_memory()
        Vd = torch.empty_like(V)
        scorer = topkrec.BatchScorer(NI, D, k, nb, engine=eng, device=dev)     # streams, double buffers, workspace
        oi = (torch.empty((Ke * nb, k), dtype=torch.int32).pin_memory(), torch.empty((Ke * nb, k), dtype=torch.float32).pin_memory())
        scorer.run(Uh[:2 * nb], V, out=(oi[0][:2 * nb], oi[1][:2 * nb]))                      # warm-up
        barrier()
        passes = 4
        t0 = time.perf_counter()
        for _ in range(passes):
            Vd.copy_(Vh, non_blocking=True)
            scorer.run(Uh, Vd, out=oi)                                                       # synchronises before returning
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": nb * Ke * passes / dt, "unit": "users/s", "h2d_bytes_per_step": nb * D * 4 + NI * D * 4 // Ke,
                      "d2h_bytes_per_step": nb * k * 8, "timed_seconds": dt,
                      "api": "topkrec.BatchScorer (evaluate.py flow): V copied once per pass of %d batches, U batches up / lists down "
                             "on side streams; %d passes" % (Ke, passes)}
        if rank == 0 and not args.skip_cpu:
            nsamp = 64
            Un, Vn = Ub[0][:nsamp].cpu().numpy(), V.cpu().numpy()
            t0 = time.perf_counter()
            S = np.dot(Un, Vn.T); np.argsort(S, axis=1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nsamp / dt, "unit": "users/s", "cores": host_threads(), "kind": "reference",
                                   "sample": "np.dot + np.argsort (evaluate.py:78,81 verbatim) on %d users x %d items" % (nsamp, NI)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="triples per mini-batch per GPU")
    ap.add_argument("--inner", type=int, default=256, help="mini-batches per bench step (20 steps x 256 x 0.46 ms = 2.3 s timed)")
    ap.add_argument("--e2e-inner", type=int, default=32, help="mini-batches per bench step of the host-buffer (e2e) measurement")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="N > 1: fused peer-memory exchange (default) or NCCL all-reduce")
    ap.add_argument("--score-users", type=int, default=18944)      # 148 tiles of 128 users: one CTA per SM
    ap.add_argument("--score-engine", default="tc", choices=["tc", "exact"])
    ap.add_argument("--score-items", type=int, default=1 << 20)
    ap.add_argument("--score-steps", type=int, default=5)
    ap.add_argument("--score-dist", default="ring", choices=["ring", "exchange"],
                    help="N > 1: ring of sweep segments (RingScorer, default) or independent shard lists + peer exchange (ShardedScorer)")
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
