#!/usr/bin/env python
"""Benchmark of the B200-native RANSAC hot path (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N ranks for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm (numpy oracle port) on host cores

Workload ("step") = one batched F-matrix RANSAC call over ``--pairs-per-step`` synthetic image pairs of the BASELINE
config-5 shape (50 000 correspondences x 8 192 hypotheses, 30 % outliers) per GPU: hypotheses solved (8-point), every
hypothesis scored against every correspondence, best hypothesis + inlier mask selected.  Pairs are independent, so N
GPUs process N disjoint pair sets with no data-path collective (weak scaling).  metric = hypothesis x correspondence
evaluations per second, whole job.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_F_EVAL = 30.0       # 12 FMA + 6 other flops per evaluation (algorithmic)
FLOP_PER_PNP_EVAL = 29.0
THR2_PNP = (1.5 / 3217.0) ** 2


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on all host cores
# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    p1, p2, idx, thr = args
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(1)
    except Exception:
        ctx = None
    from oracle import f_path as orc
    F = orc.solve_hypotheses(p1, p2, idx)
    counts = orc.score_hypotheses(F, p1, p2, thr)
    del ctx
    return counts


_POOL = None


def _cpu_pool(cores: int):
    global _POOL
    if _POOL is None and cores > 1:
        import multiprocessing as mp
        _POOL = mp.get_context("fork").Pool(cores)
        _POOL.map(abs, range(cores))                      # spin the workers up outside any timed region
    return _POOL


def cpu_reference_sample(n_points: int, n_hyp: int, cores: int, seed: int = 0):
    """Times fun.py:303-317 semantics (solve + score + threshold + count, oracle port) for ``n_hyp`` hypotheses on one
    synthetic pair of ``n_points`` correspondences, hypotheses spread over ``cores`` processes.  Returns evals/s."""
    from tsbb15_b200 import sampling, synth
    pts, _ = synth.two_view(n_points, seed=1000 + seed)
    idx = sampling.fast(n_points, n_hyp, 8, seed=seed)
    p1, p2 = pts[:, :2].T.copy(), pts[:, 2:].T.copy()
    chunks = [c for c in np.array_split(idx, cores) if len(c)]
    pool = _cpu_pool(cores)
    t0 = time.perf_counter()
    if pool is not None:
        out = pool.map(_cpu_worker, [(p1, p2, c, 1.5) for c in chunks])
    else:
        out = [_cpu_worker((p1, p2, chunks[0], 1.5))]
    dt = time.perf_counter() - t0
    best = int(np.max(np.concatenate(out)))
    return n_points * n_hyp / dt, dt, best


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_hyp = max(cores * 8, 256)                       # bounded sample: ~6 ms of numpy per hypothesis at N = 50 000
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(args.n, max(cores, 32), cores)
    times, evals = [], 0
    for s in range(args.steps):
        v, dt, _ = cpu_reference_sample(args.n, n_hyp, cores, seed=s)
        times.append(dt)
        evals += args.n * n_hyp
    total = sum(times)
    value = evals / total
    line = {
        "impl": "reference", "metric": "f_ransac_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic multi-pair F-RANSAC, config-5 pair shape (N=%d corr.), CPU sample of %d "
                               "hypotheses per step" % (args.n, n_hyp), "n_corr": args.n, "hyp_per_step": n_hyp},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d hypotheses x %d correspondences, numpy oracle port of fun.py:303-317 "
                                   "(lab3.fmatrix_stls + fmatrix_residuals), one process per core, BLAS threads=1"
                                   % (args.steps, n_hyp, args.n)},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.samples = []
        self.proc = None
        self.thread = None
        self.dev = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        rows = []
        for ts, line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((ts, float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        inside = [r for r in rows if any(a <= r[0] <= b for a, b in windows)] or rows
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in inside for k in range(4) if r[4][k].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": inside[0][2],
                "power_w_max": max(r[3] for r in inside), "reasons": reasons, "samples": len(inside)}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import tsbb15_b200 as rg
    from tsbb15_b200 import _cabi as cabi, runtime as rt, sampling, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = cabi.load_library()
    ctx = cabi.context(local)
    vp = C.c_void_p
    stream = torch.cuda.current_stream().cuda_stream

    peaks = rt.microbench(device=local)               # measured FFMA / FFMA2 / DFMA pipe rates of THIS gpu
    fp32_peak_tflops = 2.0 * max(peaks["ffma_gfma_s"], peaks["ffma2_gfma_s"]) * 1e-3

    P, N, H = args.pairs_per_step, args.n, args.hyp
    pool = max(P, (args.pool // P) * P)
    # config 5: pair p uses seed 1000 + p; ranks take disjoint pair ids
    pairs = synth.multi_pair(pool, N, first_pair=rank * pool)
    idxs = [sampling.fast(N, H, 8, seed=7919 * (rank * pool + p) + 1) for p in range(pool)]
    h_pts = torch.from_numpy(np.stack(pairs)).pin_memory()                 # (pool, N, 4) f64, pinned
    h_idx = torch.from_numpy(np.stack(idxs)).pin_memory()                  # (pool, H, 8) i32, pinned
    d_pts = h_pts.to(dev)
    d_idx = h_idx.to(dev)
    pair_off = (np.arange(P + 1, dtype=np.int32) * N)
    hyp_off = (np.arange(P + 1, dtype=np.int32) * H)
    po = pair_off.ctypes.data_as(C.POINTER(C.c_int32))
    ho = hyp_off.ctypes.data_as(C.POINTER(C.c_int32))
    d_best_idx = torch.empty(P, dtype=torch.int32, device=dev)
    d_best_cnt = torch.empty(P, dtype=torch.int32, device=dev)
    d_best_F = torch.empty(P, 9, dtype=torch.float64, device=dev)
    d_mask = torch.empty(P * N, dtype=torch.uint8, device=dev)
    h_best_idx = torch.empty(P, dtype=torch.int32).pin_memory()
    h_best_cnt = torch.empty(P, dtype=torch.int32).pin_memory()
    h_best_F = torch.empty(P, 9, dtype=torch.float64).pin_memory()
    h_mask = torch.empty(P * N, dtype=torch.uint8).pin_memory()
    n_groups = pool // P

    def step_dev(s):
        g = s % n_groups
        cabi.check(lib.rg_f_ransac_dev(vp(ctx), vp(stream), P, vp(d_pts[g * P].data_ptr()), po,
                                       vp(d_idx[g * P].data_ptr()), ho, 1.5, rg.MODE_EPI_MAX, rg.TIE_FIRST, args.solver,
                                       rg.SCORE_FP32_GUARDED, vp(d_best_idx.data_ptr()), vp(d_best_cnt.data_ptr()),
                                       vp(d_best_F.data_ptr()), vp(d_mask.data_ptr())))

    def step_host(s):
        g = s % n_groups
        cabi.check(lib.rg_f_ransac_host(vp(ctx), vp(stream), P, vp(h_pts[g * P].data_ptr()), po,
                                        vp(h_idx[g * P].data_ptr()), ho, 1.5, rg.MODE_EPI_MAX, rg.TIE_FIRST, args.solver,
                                        rg.SCORE_FP32_GUARDED, vp(h_best_idx.data_ptr()), vp(h_best_cnt.data_ptr()),
                                        vp(h_best_F.data_ptr()), vp(h_mask.data_ptr()), None, None, None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        e1.synchronize()
        w1 = time.time()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, (w0, w1)

    if args.host_slices:
        rt.set_option(2, args.host_slices, device=local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        step_dev(s)
        step_host(s)
    torch.cuda.synchronize()
    launches_per_host_call = rt.last_stats(device=local)["launches"]      # last warm-up call was step_host
    step_dev(0)
    launches_per_call = rt.last_stats(device=local, stream=stream)["launches"]

    rt.set_option(1, 1, device=local)                                       # phase events on the launching stream
    ms_dev, win_dev = timed(step_dev, args.steps)
    prof = rt.profile(device=local, stream=stream)
    rt.set_option(1, 0, device=local)
    stats = rt.last_stats(device=local, stream=stream)
    ms_e2e, win_e2e = timed(step_host, args.steps)
    # sanity: the last e2e step really produced winners
    assert int(h_best_cnt.min()) > 0.5 * 0.7 * N, "benchmark produced implausible consensus sets"

    evals_per_step = float(P) * N * H
    value = world * evals_per_step * args.steps / (ms_dev * 1e-3)
    e2e = world * evals_per_step * args.steps / (ms_e2e * 1e-3)
    score_ms = prof["score_ms"] / max(prof["calls"], 1)
    achieved_tflops = FLOP_PER_F_EVAL * evals_per_step / (score_ms * 1e-3) * 1e-12 if score_ms > 0 else None
    alg_bytes = P * (16.0 * N + 48.0 * H + 4.0 * H)
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass

    extras = {}
    cpu_baseline = None
    split = None
    if world > 1 and args.split_hypotheses:             # opt-in: a failing rank would leave the others in a collective
        try:
            split = run_split_hypotheses(rg, cabi, lib, ctx, stream, dev, torch, dist, rank, world)
        except Exception as e:                          # never take the headline down with an extra
            split = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_extras:
        extras = run_extras(args, rg, rt, cabi, lib, ctx, stream, dev, torch, fp32_peak_tflops)
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_hyp = max(cores * 48, 1024)
        v, dt, _ = cpu_reference_sample(N, n_hyp, cores)
        cpu_baseline = {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                        "sample": "1 pair x %d hypotheses x %d correspondences (%.1f s wall), numpy oracle port of "
                                  "fun.py:303-317, one process per core, BLAS threads=1" % (n_hyp, N, dt)}
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary([win_dev, win_e2e])
        line = {
            "metric": "f_ransac_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 scoring with f64 guard-band recheck; f64 solve",
            "data": "synthetic",
            "config": {"workload": "synthetic multi-pair F-RANSAC (BASELINE config 5 pair shape), %d pairs/step/GPU x "
                                   "%d correspondences x %d hypotheses, 30%% outliers, thr 1.5 px" % (P, N, H),
                       "pairs_per_step_per_gpu": P, "n_corr": N, "n_hyp": H, "parallelism": "pair-sharded x%d" % world,
                       "l2": "inputs cycle through a %d-pair pool (%.0f MB) larger than the 126 MB L2"
                             % (pool, pool * (N * 32 + H * 32) / 1e6),
                       "solver": "qr" if args.solver == 0 else "jacobi"},
            "e2e": {"value": e2e, "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": P * (N * 32 + H * 32), "d2h_bytes_per_step": P * (4 + 4 + 72 + N)},
            "gpu_launches": int(launches_per_call) * args.steps,
            "gpu_launches_e2e_region": int(launches_per_host_call) * args.steps,
            "roofline": {"bound": "fp32_ffma", "kernel": "score_packed<EpiPolicy>", "achieved": achieved_tflops,
                         "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                         "frac": (achieved_tflops / fp32_peak_tflops) if achieved_tflops else None,
                         "traffic": traffic, "kernel_ms_per_launch": score_ms,
                         "peak_source": "FFMA/FFMA2 chain micro-benchmark run on this GPU at start of bench.py "
                                        "(MEASURED_PEAKS.json has no FP32 figure)",
                         "algorithmic_flop_per_eval": FLOP_PER_F_EVAL,
                         "executed_fp32_lane_ops_per_eval": 17,
                         "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                                 "achieved_gbs": alg_bytes / (score_ms * 1e-3) * 1e-9 if score_ms > 0 else None,
                                 "peak_gbs": hbm_peak}},
            "phases_ms_per_step": {k: prof[k] / max(prof["calls"], 1) for k in
                                   ("prepare_ms", "solve_ms", "score_ms", "fixup_ms", "select_ms")},
            "guard_band": {"band_eval_fraction": stats["band_evals"] / evals_per_step,
                           "flips_per_step": stats["flips"], "overflow": stats["overflow"]},
            "pipes": peaks, "clocks": clocks, "cpu_baseline": cpu_baseline,
        }
        line.update(extras)
        if split is not None:
            line["config3_split_hypotheses"] = split
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_extras(args, rg, rt, cabi, lib, ctx, stream, dev, torch, fp32_peak) -> dict:
    """Short, separately timed runs of the other BASELINE configs (reported, not the headline)."""
    from tsbb15_b200 import sampling, synth
    vp = C.c_void_p
    out = {}

    def time_dev(fn, reps):
        fn(); torch.cuda.synchronize()
        rt.set_option(1, 1)                      # phase events from here on (the warm-up call is not profiled)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- config 3: single pair 100k x 16k, threshold sweep, both criteria -------------------------------------
    N3, H3 = 100000, 16384
    pts, _ = synth.two_view(N3, seed=1)
    idx = sampling.fast(N3, H3, 8, seed=2)
    d_pts = torch.from_numpy(pts).to(dev); d_idx = torch.from_numpy(idx).to(dev)
    po = (np.array([0, N3], dtype=np.int32)); ho = np.array([0, H3], dtype=np.int32)
    ob = torch.empty(2, dtype=torch.int32, device=dev); oF = torch.empty(9, dtype=torch.float64, device=dev)
    om = torch.empty(N3, dtype=torch.uint8, device=dev)
    sweep = {}
    for mode, name in ((rg.MODE_EPI_MAX, "epi_max"), (rg.MODE_SAMPSON, "sampson")):
        for thr in ((0.25, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0) if mode == rg.MODE_EPI_MAX else (1.5,)):
            def call():
                cabi.check(lib.rg_f_ransac_dev(vp(ctx), vp(stream), 1, vp(d_pts.data_ptr()),
                                               po.ctypes.data_as(C.POINTER(C.c_int32)), vp(d_idx.data_ptr()),
                                               ho.ctypes.data_as(C.POINTER(C.c_int32)), float(thr), mode, rg.TIE_FIRST,
                                               args.solver, rg.SCORE_FP32_GUARDED, vp(ob.data_ptr()),
                                               vp(ob[1:].data_ptr()), vp(oF.data_ptr()), vp(om.data_ptr())))
            ms = time_dev(call, 5)
            pr = rt.profile(stream=stream); rt.set_option(1, 0)
            st = rt.last_stats(stream=stream)
            sc = pr["score_ms"] / max(pr["calls"], 1)
            sweep["%s_thr%g" % (name, thr)] = {
                "ms": ms, "evals_per_s": N3 * H3 / (ms * 1e-3), "score_kernel_ms": sc,
                "score_kernel_frac_of_fp32_peak": FLOP_PER_F_EVAL * N3 * H3 / (sc * 1e-3) * 1e-12 / fp32_peak,
                "best_count": int(ob[1].item()), "band_eval_fraction": st["band_evals"] / (N3 * H3)}
    out["config3_single_pair_100k_x_16k"] = sweep

    # ---- config 4: PnP 1M x 8192 -------------------------------------------------------------------------------
    N4, H4 = 1000000, 8192
    X, y, _ = synth.pnp_scene(N4, seed=4)
    pidx = sampling.fast(N4, H4, 6, seed=2)
    dX = torch.from_numpy(X).to(dev); dy = torch.from_numpy(y).to(dev); dI = torch.from_numpy(pidx).to(dev)
    oRt = torch.empty(12, dtype=torch.float64, device=dev); omp = torch.empty(N4, dtype=torch.uint8, device=dev)

    def pcall():
        cabi.check(lib.rg_pnp_ransac_dev(vp(ctx), vp(stream), N4, N4, vp(dX.data_ptr()), vp(dy.data_ptr()), H4, 6,
                                         vp(dI.data_ptr()), THR2_PNP, rg.SCORE_FP32_GUARDED, vp(ob.data_ptr()),
                                         vp(ob[1:].data_ptr()), vp(oRt.data_ptr()), vp(omp.data_ptr())))
    ms = time_dev(pcall, 3)
    pr = rt.profile(stream=stream); rt.set_option(1, 0)
    st = rt.last_stats(stream=stream)
    sc = pr["score_ms"] / max(pr["calls"], 1)
    out["config4_pnp_1M_x_8192"] = {
        "ms": ms, "poses_per_s": H4 / (ms * 1e-3), "evals_per_s": float(N4) * H4 / (ms * 1e-3),
        "solve_ms": pr["solve_ms"] / max(pr["calls"], 1), "score_kernel_ms": sc,
        "score_kernel_frac_of_fp32_peak": FLOP_PER_PNP_EVAL * N4 * H4 / (sc * 1e-3) * 1e-12 / fp32_peak,
        "best_count": int(ob[1].item()), "band_eval_fraction": st["band_evals"] / (float(N4) * H4)}

    # ---- config 2: Dino sequence through the host API (real data shapes, launch/latency bound) ---------------------
    try:
        pairs = [np.ascontiguousarray(np.hstack(synth.dino_noisy_pair(i, i + 1))) for i in range(35)]
        idl = sampling.fast_batch([p.shape[0] for p in pairs], 10000, 8, seed=0)     # one buffer: no host concatenation
        rt.f_ransac_batched(pairs, idl, thr=1.5)
        t0 = time.perf_counter()
        for _ in range(5):
            r = rt.f_ransac_batched(pairs, idl, thr=1.5)
        dt = (time.perf_counter() - t0) / 5
        ev = sum(p.shape[0] for p in pairs) * 10000.0
        views = [synth.dino_view_2d3d(i) for i in range(36)]
        vidx = [sampling.fast(v[0].shape[0], 1024, 6, seed=i) for i, v in enumerate(views)]
        Xl, yl = [v[0] for v in views], [v[1] for v in views]
        rp = rt.pnp_ransac_batched(Xl, yl, vidx, THR2_PNP)
        t0 = time.perf_counter()
        for _ in range(5):
            rp = rt.pnp_ransac_batched(Xl, yl, vidx, THR2_PNP)
        dtp = (time.perf_counter() - t0) / 5
        out["config2_dino_sequence"] = {
            "f_35_pairs_x_10000_hyp_host_call_ms": dt * 1e3, "f_evals_per_s_e2e": ev / dt,
            "f_inlier_counts": [int(c) for c in r["best_count"][:5]],
            "pnp_36_views_x_1024_hyp_one_host_call_ms": dtp * 1e3, "pnp_poses_per_s_e2e": 36 * 1024 / dtp,
            "pnp_consensus": [int(c) for c in rp["best_count"][:5]], "pnp_view_sizes": [int(v[0].shape[0]) for v in views[:5]]}
    except Exception as e:                                            # fixture missing: report, do not fail the bench
        out["config2_dino_sequence"] = {"error": repr(e)}

    # ---- rows next to the hot path: triangulation, relative pose, 2D<->3D matching -----
    try:
        out["next_rows"] = run_next_rows(args, rg, rt, cabi, lib, ctx, stream, dev, torch)
    except Exception as e:
        out["next_rows"] = {"error": repr(e)}
    return out


def _tri_cpu_worker(a):
    from oracle import geom_path as og
    C1, C2, x1, x2 = a
    return og.triangulate_optimal_batch(C1, C2, x1, x2)


def run_split_hypotheses(rg, cabi, lib, ctx, stream, dev, torch, dist, rank, world) -> dict:
    """BASELINE config 3 (ONE pair, 100 000 correspondences x 16 384 hypotheses) with the hypotheses split over the ranks
    (SURVEY 8e case 2): every rank holds all correspondences, scores its block with rg_f_ransac_dev, packs
    (count, global index) into a key on the device (rg_argmax_pack_dev), ONE 8-byte NCCL max-all-reduce picks the winner,
    rg_argmax_unpack_dev decodes it; the owner's F is broadcast (72 bytes).  Device resident, CUDA events, max over ranks."""
    from tsbb15_b200 import sampling, synth
    vp, pi32 = C.c_void_p, C.POINTER(C.c_int32)
    N, H = 100000, 16384
    pts, _ = synth.two_view(N, seed=1)
    idx = sampling.fast(N, H, 8, seed=2)
    lo, hi = rank * H // world, (rank + 1) * H // world
    d_pts = torch.from_numpy(np.ascontiguousarray(pts)).to(dev)
    d_idx = torch.from_numpy(np.ascontiguousarray(idx[lo:hi])).to(dev)
    d_all = torch.from_numpy(np.ascontiguousarray(idx)).to(dev) if rank == 0 else None
    po = np.array([0, N], dtype=np.int32)
    ho = np.array([0, hi - lo], dtype=np.int32)
    d_bi = torch.empty(1, dtype=torch.int32, device=dev)
    d_bc = torch.empty(1, dtype=torch.int32, device=dev)
    d_F = torch.empty(9, dtype=torch.float64, device=dev)
    d_key = torch.empty(1, dtype=torch.int64, device=dev)
    d_gi = torch.empty(1, dtype=torch.int32, device=dev)
    d_gc = torch.empty(1, dtype=torch.int32, device=dev)
    d_Fw = torch.empty(9, dtype=torch.float64, device=dev)

    def call():
        cabi.check(lib.rg_f_ransac_dev(vp(ctx), vp(stream), 1, vp(d_pts.data_ptr()), po.ctypes.data_as(pi32), vp(d_idx.data_ptr()),
                                       ho.ctypes.data_as(pi32), 1.5, rg.MODE_EPI_MAX, rg.TIE_FIRST, rg.SOLVER_QR,
                                       rg.SCORE_FP32_GUARDED, vp(d_bi.data_ptr()), vp(d_bc.data_ptr()), vp(d_F.data_ptr()), None))
        cabi.check(lib.rg_argmax_pack_dev(vp(stream), 1, vp(d_bi.data_ptr()), vp(d_bc.data_ptr()), lo, vp(d_key.data_ptr())))
        dist.all_reduce(d_key, op=dist.ReduceOp.MAX)
        cabi.check(lib.rg_argmax_unpack_dev(vp(stream), 1, vp(d_key.data_ptr()), vp(d_gi.data_ptr()), vp(d_gc.data_ptr())))

    def ev(fn, reps):
        fn(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = ev(call, 10)
    gi, gc = int(d_gi.item()), int(d_gc.item())
    owner = next(r for r in range(world) if r * H // world <= gi < (r + 1) * H // world) if gi >= 0 else 0
    d_Fw.copy_(d_F)
    dist.broadcast(d_Fw, src=owner)                      # the winner's F from its owner
    out = {"workload": "config 3: one pair, 100000 correspondences x 16384 hypotheses split over the ranks",
           "ms": ms, "evals_per_s": float(N) * H / (ms * 1e-3), "winner": gi, "count": gc, "owner_rank": owner,
           "collective": "one 8-byte ncclAllReduce(max) per call + a 72-byte broadcast of the winner's F"}
    if rank == 0:                                        # the same problem on one GPU: identical winner
        d_bi1 = torch.empty(1, dtype=torch.int32, device=dev)
        d_bc1 = torch.empty(1, dtype=torch.int32, device=dev)
        d_F1 = torch.empty(9, dtype=torch.float64, device=dev)
        ho1 = np.array([0, H], dtype=np.int32)
        cabi.check(lib.rg_f_ransac_dev(vp(ctx), vp(stream), 1, vp(d_pts.data_ptr()), po.ctypes.data_as(pi32), vp(d_all.data_ptr()),
                                       ho1.ctypes.data_as(pi32), 1.5, rg.MODE_EPI_MAX, rg.TIE_FIRST, rg.SOLVER_QR,
                                       rg.SCORE_FP32_GUARDED, vp(d_bi1.data_ptr()), vp(d_bc1.data_ptr()), vp(d_F1.data_ptr()), None))
        torch.cuda.synchronize()
        out["same_winner_as_one_gpu"] = bool(int(d_bi1.item()) == gi and int(d_bc1.item()) == gc
                                             and torch.equal(d_F1, d_Fw))
    return out


def run_next_rows(args, rg, rt, cabi, lib, ctx, stream, dev, torch) -> dict:
    """Device-resident and end-to-end rates of the kernels either side of the RANSAC path, with the numpy oracle port of
    the reference loop timed beside them (bounded sample)."""
    vp = C.c_void_p
    pi32 = C.POINTER(C.c_int32)
    g = np.load(os.path.join(ROOT, "tests", "golden", "dino_data.npz"))
    Ps = g["Ps"]
    rng = np.random.default_rng(0)
    out = {}

    def ev_time(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / reps

    # triangulation: 1M correspondences of one camera pair, 0.5 px noise (config-4 sized point set)
    N = 1_000_000
    Xw = np.column_stack([rng.uniform(-0.045, 0.045, N), rng.uniform(-0.08, 0.03, N), rng.uniform(-0.72, -0.54, N),
                          np.ones(N)])
    proj = lambda Pm: (Xw @ Pm.T)[:, :2] / (Xw @ Pm.T)[:, 2:]
    x1 = proj(Ps[0]) + rng.normal(0, 0.5, (N, 2))
    x2 = proj(Ps[1]) + rng.normal(0, 0.5, (N, 2))
    d1, d2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
    dC1, dC2 = torch.from_numpy(Ps[0:1].copy()).to(dev), torch.from_numpy(Ps[1:2].copy()).to(dev)
    dX = torch.empty((N, 3), dtype=torch.float64, device=dev)
    off = np.array([0, N], dtype=np.int32)
    tri = {}
    for method, name in ((0, "optimal"), (1, "linear")):
        def call():
            cabi.check(lib.rg_triangulate_dev(vp(ctx), vp(stream), 1, vp(dC1.data_ptr()), vp(dC2.data_ptr()),
                                              off.ctypes.data_as(pi32), vp(d1.data_ptr()), vp(d2.data_ptr()), method,
                                              vp(dX.data_ptr())))
        ms = ev_time(call, 5)
        t0 = time.perf_counter()
        rt.triangulate(Ps[0], Ps[1], [x1], [x2], method=method)
        dt = time.perf_counter() - t0
        tri[name] = {"ms": ms, "points_per_s": N / (ms * 1e-3), "e2e_points_per_s_host_call": N / dt,
                     "algorithmic_bytes_per_point": 56, "hbm_gbs": 56.0 * N / (ms * 1e-3) * 1e-9}
    cores = os.cpu_count() or 1
    per = 48
    pool = _cpu_pool(cores)
    jobs = [(Ps[0], Ps[1], x1[k * per:(k + 1) * per], x2[k * per:(k + 1) * per]) for k in range(cores)]
    t0 = time.perf_counter()
    ref = pool.map(_tri_cpu_worker, jobs) if pool is not None else [_tri_cpu_worker(jobs[0])]
    dt = time.perf_counter() - t0
    tri["cpu_baseline"] = {"value": per * len(jobs) / dt, "unit": "points/s", "cores": cores, "kind": "port",
                           "sample": "%d points, numpy oracle port of lab3.triangulate_optimal, one process per core"
                                     % (per * len(jobs))}
    got = rt.triangulate(Ps[0], Ps[1], [x1[:per]], [x2[:per]])[0]
    tri["max_rel_err_vs_oracle_sample"] = float(np.abs(got - ref[0]).max() / np.abs(ref[0]).max())
    out["triangulate_1M_points"] = tri

    # relative pose for 4096 pairs (config-5 pair count): E of random camera pairs + one correspondence each
    from oracle import geom_path as og
    Pn = 4096
    Ks = og.camera_resectioning(Ps[0])[0]
    ii = rng.integers(0, 35, Pn)
    Es, y1s, y2s = np.empty((Pn, 3, 3)), np.empty((Pn, 2)), np.empty((Pn, 2))
    Kinv = np.linalg.inv(Ks)
    Fs = [og.fmatrix_from_cameras(Ps[i], Ps[i + 1]) for i in range(35)]
    for k, i in enumerate(ii):
        Es[k] = Ks.T @ Fs[i] @ Ks
        Xp = np.array([rng.uniform(-0.04, 0.04), rng.uniform(-0.07, 0.02), rng.uniform(-0.7, -0.56), 1.0])
        a, b = Ps[i] @ Xp, Ps[i + 1] @ Xp
        y1s[k] = (Kinv @ (a / a[2]))[:2]
        y2s[k] = (Kinv @ (b / b[2]))[:2]
    dE, dy1, dy2 = (torch.from_numpy(a).to(dev) for a in (Es, y1s, y2s))
    dRt = torch.empty((Pn, 12), dtype=torch.float64, device=dev)
    dw = torch.empty(2 * Pn, dtype=torch.int32, device=dev)

    def rcall():
        cabi.check(lib.rg_relative_pose_dev(vp(ctx), vp(stream), Pn, vp(dE.data_ptr()), None, 0, vp(dy1.data_ptr()),
                                            vp(dy2.data_ptr()), vp(dRt.data_ptr()), vp(dw.data_ptr()),
                                            vp(dw[Pn:].data_ptr())))
    ms = ev_time(rcall, 5)
    t0 = time.perf_counter()
    for k in range(32):
        og.relative_camera_pose(Es[k], y1s[k], y2s[k])
    dt = time.perf_counter() - t0
    out["relative_pose_4096_pairs"] = {"ms": ms, "pairs_per_s": Pn / (ms * 1e-3), "resolved": int((dw[:Pn] >= 0).sum().item()),
                                       "cpu_baseline": {"value": 32 / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                                        "sample": "32 pairs, numpy oracle port of fun.relative_camera_pose"}}

    # main.py:54-76 for 512 pairs x 2000 correspondences in one call (E = K^T F K, relative pose, triangulation)
    Pt, Nt = 512, 2000
    it = rng.integers(0, 35, Pt)
    tv_pts = np.empty((Pt * Nt, 4))
    tv_F = np.empty((Pt, 3, 3))
    for k, i in enumerate(it):
        Xs = np.column_stack([rng.uniform(-0.04, 0.04, Nt), rng.uniform(-0.07, 0.02, Nt), rng.uniform(-0.7, -0.56, Nt),
                              np.ones(Nt)])
        a, b = Xs @ Ps[i].T, Xs @ Ps[i + 1].T
        tv_pts[k * Nt:(k + 1) * Nt, :2] = a[:, :2] / a[:, 2:]
        tv_pts[k * Nt:(k + 1) * Nt, 2:] = b[:, :2] / b[:, 2:]
        tv_F[k] = Fs[i]
    tv_pts += rng.normal(0, 0.3, tv_pts.shape)
    d_tp, d_tF = torch.from_numpy(tv_pts).to(dev), torch.from_numpy(tv_F).to(dev)
    d_tRt = torch.empty((Pt, 12), dtype=torch.float64, device=dev)
    d_tw = torch.empty(Pt, dtype=torch.int32, device=dev)
    d_tX = torch.empty((Pt * Nt, 3), dtype=torch.float64, device=dev)
    toff = (np.arange(Pt + 1, dtype=np.int32) * Nt)
    Kc = np.ascontiguousarray(Ks)

    def tcall():
        cabi.check(lib.rg_two_view_init_dev(vp(ctx), vp(stream), Pt, vp(d_tp.data_ptr()), toff.ctypes.data_as(pi32),
                                            vp(d_tF.data_ptr()), vp(Kc.ctypes.data), None, vp(d_tRt.data_ptr()),
                                            vp(d_tw.data_ptr()), vp(d_tX.data_ptr())))
    ms = ev_time(tcall, 5)
    out["two_view_init_512_pairs_x_2000"] = {"ms": ms, "pairs_per_s": Pt / (ms * 1e-3), "points_per_s": Pt * Nt / (ms * 1e-3),
                                             "resolved": int((d_tw >= 0).sum().item()),
                                             "finite_points": int(torch.isfinite(d_tX).all(dim=1).sum().item())}

    # gold-standard refinement (second half of fun.getFFromLabCode): the reference's own case (257 inliers of the noisy
    # Dino pair (0,1); SciPy: 5482 residual evaluations, 267 s on this project's build box) and a config-5 sized batch
    try:
        gsg = np.load(os.path.join(ROOT, "tests", "golden", "gs_golden.npz"))
        gpts = np.ascontiguousarray(np.hstack([gsg["in1"].T, gsg["in2"].T]))
        rt.gold_standard([gpts], gsg["F0"][None])
        t0 = time.perf_counter()
        for _ in range(5):
            gres = rt.gold_standard([gpts], gsg["F0"][None])
        dt = (time.perf_counter() - t0) / 5
        Pg, Ng = 16, 50000
        big = []
        for k in range(Pg):
            Xs = np.column_stack([rng.uniform(-0.04, 0.04, Ng), rng.uniform(-0.07, 0.02, Ng), rng.uniform(-0.7, -0.56, Ng),
                                  np.ones(Ng)])
            a, b = Xs @ Ps[k].T, Xs @ Ps[k + 1].T
            big.append(np.hstack([a[:, :2] / a[:, 2:], b[:, :2] / b[:, 2:]]) + rng.normal(0, 0.5, (Ng, 4)))
        d_gp = torch.from_numpy(np.concatenate(big)).to(dev)
        d_gF = torch.from_numpy(np.stack([Fs[k] for k in range(Pg)])).to(dev)
        d_gout = torch.empty((Pg, 9), dtype=torch.float64, device=dev)
        d_gcost = torch.empty(Pg, dtype=torch.float64, device=dev)
        d_git = torch.empty(2 * Pg, dtype=torch.int32, device=dev)
        goff = (np.arange(Pg + 1, dtype=np.int32) * Ng)

        def gcall():
            cabi.check(lib.rg_gold_standard_dev(vp(ctx), vp(stream), Pg, vp(d_gp.data_ptr()), goff.ctypes.data_as(pi32),
                                                vp(d_gF.data_ptr()), None, 50, 1e-12, vp(d_gout.data_ptr()),
                                                vp(d_gcost.data_ptr()), vp(d_git.data_ptr()), vp(d_git[Pg:].data_ptr()), None))
        msg = ev_time(gcall, 3)
        out["gold_standard"] = {
            "dino_noisy_pair_257_inliers": {"host_call_ms": dt * 1e3, "cost": float(gres["cost"][0]),
                                            "iters": int(gres["iters"][0]), "status": int(gres["status"][0]),
                                            "reference_scipy": {"cost": float(gsg["scipy_cost"]), "nfev": int(gsg["scipy_nfev"]),
                                                                "seconds_on_build_box": 267}},
            "batch_16_pairs_x_50000": {"ms": msg, "points_per_s": Pg * Ng / (msg * 1e-3),
                                       "iters": [int(v) for v in d_git[:Pg].cpu().numpy()[:4]],
                                       "status": [int(v) for v in d_git[Pg:].cpu().numpy()[:4]],
                                       "cost_per_point": float(d_gcost.mean().item()) / Ng}}
    except Exception as e:
        out["gold_standard"] = {"error": repr(e)}

    # bundle adjustment (Tables.BundleAdjustment2): the 36-view Dino scene the golden file holds (reference: SciPy trf with
    # a finite-difference Jacobian, 14 residual sweeps, ~8 s) and the oracle's SciPy call timed here on the 8-view scene
    try:
        from oracle import ba_path as oba
        bag = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
        bsc = lambda nv: tuple(bag[f"v{nv}_" + k] for k in ("cams0", "pts0", "uv", "cam_idx", "pt_idx"))
        b36 = bsc(36)
        rt.bundle_adjust(*b36, ftol=1e-4)
        t0 = time.perf_counter()
        for _ in range(5):
            bres = rt.bundle_adjust(*b36, ftol=1e-4)
        dt = (time.perf_counter() - t0) / 5
        d_bc = torch.from_numpy(b36[0].copy()).to(dev)
        d_bp = torch.from_numpy(b36[1].copy()).to(dev)
        d_buv = torch.from_numpy(b36[2].copy()).to(dev)
        d_bout = torch.empty(1, dtype=torch.float64, device=dev)
        d_bit = torch.empty(2, dtype=torch.int32, device=dev)
        bci = np.ascontiguousarray(b36[3], dtype=np.int32)
        bpi = np.ascontiguousarray(b36[4], dtype=np.int32)
        n_it = 8                                     # fixed number of LM iterations from the same start: per-iteration device time

        def bcall():
            d_bc.copy_(torch.from_numpy(b36[0]), non_blocking=False)
            d_bp.copy_(torch.from_numpy(b36[1]), non_blocking=False)
            cabi.check(lib.rg_bundle_adjust_dev(vp(ctx), vp(stream), 36, b36[1].shape[0], b36[2].shape[0], vp(d_bc.data_ptr()),
                                                vp(d_bp.data_ptr()), vp(d_buv.data_ptr()), bci.ctypes.data_as(pi32),
                                                bpi.ctypes.data_as(pi32), 1, n_it, 0.0, vp(d_bout.data_ptr()),
                                                vp(d_bit.data_ptr()), vp(d_bit[1:].data_ptr())))
        msb = ev_time(bcall, 3)
        b8 = bsc(8)
        t0 = time.perf_counter()
        _, _, sol8 = oba.bundle_adjust_scipy(*b8)
        dt8 = time.perf_counter() - t0
        r8 = rt.bundle_adjust(*b8, ftol=1e-4)
        t0 = time.perf_counter()
        for _ in range(5):
            r8 = rt.bundle_adjust(*b8, ftol=1e-4)
        dt8g = (time.perf_counter() - t0) / 5
        from tsbb15_b200 import synth as _synth
        big = _synth.ba_scene(36, 20000, track=8)
        rbig = rt.bundle_adjust(*big[:5], ftol=1e-6)
        t0 = time.perf_counter()
        rbig = rt.bundle_adjust(*big[:5], ftol=1e-6)
        dtbig = time.perf_counter() - t0
        out["bundle_adjust"] = {
            "synthetic_36_views_20000_points_160000_obs": {"host_call_ms": dtbig * 1e3, "cost": rbig["cost"],
                                                           "iters": rbig["iters"], "status": rbig["status"],
                                                           "observations_per_s_per_iteration":
                                                               160000 * rbig["iters"] / dtbig},
            "dino_36_views_676_points_4165_obs": {
                "host_call_ms": dt * 1e3, "cost": bres["cost"], "iters": bres["iters"], "status": bres["status"],
                "device_ms_per_lm_iteration": msb / n_it,
                "reference_scipy": {"cost": float(bag["v36_scipy_cost"]), "nfev": int(bag["v36_scipy_nfev"]),
                                    "seconds_on_build_box": 8.1}},
            "dino_8_views_113_points_560_obs": {
                "host_call_ms": dt8g * 1e3, "cost": r8["cost"], "iters": r8["iters"],
                "cpu_baseline": {"value": dt8 * 1e3, "unit": "ms per bundle adjustment", "cores": 1, "kind": "port",
                                 "cost": float(sol8.cost), "nfev": int(sol8.nfev),
                                 "sample": "the reference's SciPy call (tables.py:317) on the oracle's vectorised EpsilonBA, "
                                           "whole problem"}}}
    except Exception as e:
        out["bundle_adjust"] = {"error": repr(e)}

    # 2D<->3D match loop of Tables.addNewView: 20 000 queries against 20 000 observations, 2/3 of them present
    M = Nq = 20000
    obs = np.column_stack([rng.uniform(-0.1, 0.1, (M, 2)), np.ones(M)])
    q = np.column_stack([rng.uniform(-0.1, 0.1, (Nq, 2)), np.ones(Nq)])
    take = rng.uniform(size=Nq) < 0.67
    q[take] = obs[rng.integers(0, M, int(take.sum()))]
    dobs, dq = torch.from_numpy(obs).to(dev), torch.from_numpy(q).to(dev)
    dm = torch.empty(Nq, dtype=torch.int32, device=dev)

    def mcall():
        cabi.check(lib.rg_match_first_within_dev(vp(ctx), vp(stream), 3, M, vp(dobs.data_ptr()), Nq, vp(dq.data_ptr()),
                                                 1e-4, vp(dm.data_ptr())))
    ms = ev_time(mcall, 5)
    t0 = time.perf_counter()
    refm = og.match_first_within(obs, q[:24], 1e-4)
    dt = time.perf_counter() - t0
    out["match_20000_x_20000"] = {"ms": ms, "queries_per_s": Nq / (ms * 1e-3),
                                  "matched": int((dm >= 0).sum().item()),
                                  "bit_exact_vs_oracle_sample": bool(np.array_equal(dm[:24].cpu().numpy(), refm)),
                                  "cpu_baseline": {"value": 24 / dt, "unit": "queries/s", "cores": 1, "kind": "port",
                                                   "sample": "24 queries x 20000 observations, the literal loop of "
                                                             "tables.py:116-124"}}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=16)
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--hyp", type=int, default=8192)
    ap.add_argument("--pool", type=int, default=80)
    ap.add_argument("--solver", type=int, default=0)
    ap.add_argument("--host-slices", type=int, default=0, help="sub-batches of the host entry point (0 = automatic)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--split-hypotheses", action="store_true",
                    help="under torchrun: also time config 3 with one pair's hypotheses split over the ranks (NCCL argmax)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_ours(args)
    finally:
        if _POOL is not None:                       # CPU-baseline workers: shut down before interpreter teardown
            _POOL.close()
            _POOL.join()


if __name__ == "__main__":
    main()
