#!/usr/bin/env python
"""bench.py - Poisson SpMM iterations/sec on the 70k-node k=10 graph with 10 classes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--iters I]

One "step" = I iterations (default 1000) of u <- Db + P u on the device-resident graph = ONE launch of the
persistent kernel.  `value` is whole-job iterations/s with all inputs resident in HBM; `e2e` is the same
metric through the reference-facing API gl.ssl.poisson(...).fit(...) with HOST buffers (host<->device copies,
graph normalisation and the result read-back inside the timed region).  N>1: the 70k graph fits one GPU, so
ranks run independent label sets on replicas of the graph (the reference's own ssl_trials parallelism,
ssl.py:390-396) with no data-path collective - weak scaling.  Timing: CUDA events on the launching stream
per step, L2 flushed between steps, max over ranks.

--impl reference times the reference's CPU implementation of the same loop (scipy csr_matvecs through the
oracle restatement, ssl.py:667-669) on the host, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES, K_NN, N_CLASSES = 70000, 10, 10
METRIC = "Poisson SpMM iterations/sec on 70k-node k=10 graph, 10 classes"
UNIT = "iterations/s"


def build_workload(seed=0):
    """The synthetic 70k-node graph BASELINE.md section 2 measured the reference on ("Poisson GD iterate, synthetic":
    70 000 points, 10 Gaussian blobs in R^8, k=10, nnz = 1 004 292 - within 1 % of the real MNIST graph's 1 014 572
    that section 3 quotes the roofline on), built with scipy's cKDTree exactly as the reference does for low d
    (~10 s, untimed setup, identical for both arms).  The graph the GPU kNN search builds from 128-d features is
    measured too (`other_rows`): isotropic 128-d Gaussians give hub nodes with > 1000 neighbours, which real
    kNN graphs (MNIST: max 45) do not have."""
    from oracle import gl_oracle as orc
    from scipy import sparse
    X, labels = orc.synthetic_blobs(N_NODES, 8, c=N_CLASSES, seed=seed)
    ind, dist = orc.knnsearch(X.astype(np.float64), K_NN + 1, method="kdtree")
    W = orc.knn_weights(ind, dist, K_NN)
    return sparse.csr_matrix(W), labels


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (ncu --set full,
# profiles/r1_dataflow_v2_nopoll_ncu_details.txt / r1_barrier_ncu_details.txt)
NCU_DRAM_BYTES_PER_LAUNCH = {"dataflow": 18776064 + 1025024, "barrier": 17295360 + 296704, "step": None}


def other_rows(W, labels, ti):
    """Short measurements of the other rows of SURVEY.md 8(a) on this box (rank 0, untimed part of the bench line):
    kNN build at config-2 size, the iterate on the 128-d (hub-heavy) graph that search produces, Laplace CG."""
    import torch
    import graphlearning_b200 as gl
    from graphlearning_b200 import device as gdev, knn_gpu
    from oracle import gl_oracle as orc
    out = {}
    try:
        X, lab = orc.synthetic_blobs(N_NODES, 128, c=N_CLASSES, seed=0)
        X = X.astype(np.float64)
        knn_gpu.knnsearch_gpu(X[:4096], K_NN + 1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ind, dist_ = knn_gpu.knnsearch_gpu(X, K_NN + 1)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        out["knn_70k_x_128_k10"] = {"seconds_host_to_host": t, "fp32_TFLOPs_on_2n2d": 2.0 * N_NODES * N_NODES * 128 / t / 1e12,
                                    "fallback_rows": knn_gpu.last_stats.get("fallback_rows"),
                                    "reference_cpu_seconds": "2460 (cKDTree, 1 thread, BASELINE.md section 2)"}
        Wh = orc.knn_weights(ind, dist_, K_NN)
        deg = np.diff(Wh.indptr)
        th = orc.one_per_class(lab, rate=1, seed=0)
        op = gdev.PoissonOperator(Wh)
        Db = op.source_to_Db(orc.poisson_source(N_NODES, th, lab[th])[0])
        u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
        best = 1e30
        for _ in range(3):
            u0.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); op.iterate(Db, 500, u0, u1); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        b = algorithmic_bytes(N_NODES, Wh.nnz, N_CLASSES)
        out["poisson_on_128d_graph"] = {"nnz": int(Wh.nnz), "max_row": int(deg.max()), "p99_row": int(np.percentile(deg, 99)),
                                        "kernel": op.kind(N_CLASSES), "gate_every": op.gate(N_CLASSES), "us_per_iteration": best * 1e3 / 500,
                                        "iterations_per_s": 500 / (best * 1e-3), "achieved_GBs": b * 500 / (best * 1e-3) / 1e9}
        m = gl.ssl.laplace(W)
        t5 = orc.one_per_class(labels, rate=5, seed=0)
        m.fit(t5, labels[t5])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(t5, labels[t5])
        t = time.perf_counter() - t0
        out["laplace_cg_fit"] = {"seconds": t, "cg_iterations": int(m.iterations), "gpu_launches": int(m.gpu_launches),
                                 "note": "gl.ssl.laplace(W).fit, 5 labels/class, tol 1e-5; includes the scipy system assembly"}
        # config 4: 50 eigenpairs of the normalised Laplacian (graph.eigen_decomp on the block kernels of spectral.cu)
        G = gl.graph(W)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vals, vecs = G.eigen_decomp(normalization="normalized", k=50)
        t = time.perf_counter() - t0
        out["eigen_decomp_k50"] = {"seconds": t, "spmm_launches": int(G.eigen_info["spmm"]), "residual": float(G.eigen_info["residual"]),
                                   "lambda_50": float(vals[-1]), "reference_cpu_seconds": "6.1 (ARPACK svds, BASELINE.md section 2)"}
        # p-Laplace / AMLE sweeps (bit-exact Gauss-Seidel / Jacobi of c_code/lp_iterate.cpp), one class against the rest
        val = (labels[t5] == 0).astype(np.float64)
        G.plaplace(t5, val, 3, max_num_it=30)
        for name, fn in (("plaplace_p3_fast", lambda: G.plaplace(t5, val, 3)), ("amle_weighted", lambda: G.amle(t5, val, tol=1e-3, max_num_it=300))):
            t0 = time.perf_counter(); fn(); t = time.perf_counter() - t0
            out[name] = {"seconds_host_to_host": t, "sweeps": int(G.sweeps), "us_per_sweep": 1e6 * t / max(1, G.sweeps)}
        mp = gl.ssl.plaplace(W, p=3)                                # one-vs-rest over the 10 classes, batched sweep kernel
        t0 = time.perf_counter(); mp.fit(t5, labels[t5]); t = time.perf_counter() - t0
        out["ssl_plaplace_p3_fit_10_classes"] = {"seconds": t, "sweeps_per_class": [int(x) for x in mp.graph.sweeps],
                                                 "reference_cpu_seconds": "15.7 (BASELINE.md section 2, MNIST-size graph)"}
    except Exception as e:                                   # never lose the headline line over an extra
        out["error"] = repr(e)
    return out


def algorithmic_bytes(n, nnz, c):
    """fp32 values + int32 columns, int32 row pointers, read u, read Db, write u."""
    return nnz * 8 + (n + 1) * 4 + 3 * n * c * 4


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_loop(W, labels, train_ind, iters):
    """The reference's CPU loop (ssl.py:667-669) on the oracle's setup: returns seconds for `iters` iterations."""
    from oracle import gl_oracle as orc
    s = orc.poisson_gd_setup(W, train_ind, labels[train_ind])
    P, Db, RW, v = s["P"], s["Db"], s["RW"], s["v"]
    u = np.zeros_like(Db)
    t0 = time.perf_counter()
    for _ in range(iters):
        u = Db + P * u
        v = RW * v
    return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import gl_oracle as orc
    W, labels = build_workload()
    ti = orc.one_per_class(labels, rate=1, seed=0)
    iters = args.ref_iters
    for _ in range(args.warmup):
        cpu_reference_loop(W, labels, ti, 2)
    times = [cpu_reference_loop(W, labels, ti, iters) for _ in range(args.steps)]
    total = float(np.sum(times))
    value = iters * args.steps / total
    sample = "%d iterations of u=Db+P*u; v=RW*v per step on the full 70k graph (scipy csr_matvecs, fp64)" % iters
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: 70k nodes, k=10 kNN graph (10 Gaussian blobs), 10 classes, 1 label/class",
                   "n": int(W.shape[0]), "nnz": int(W.nnz), "classes": N_CLASSES, "iterations_per_step": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import graphlearning_b200 as gl
    from graphlearning_b200 import device as gdev
    from oracle import gl_oracle as orc

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W, labels = build_workload()
    n, nnz, c = W.shape[0], W.nnz, N_CLASSES
    # every rank works on its own label set (replicas of the graph, independent trials)
    ti = orc.one_per_class(labels, rate=1, seed=rank)
    source = orc.poisson_source(n, ti, labels[ti])[0]
    iters = args.iters

    op = gdev.PoissonOperator(W, kind=os.environ.get("GLB_BENCH_KIND", "auto"))   # profiler runs pin the kernel: the plan's trial timing is meaningless under ncu
    Db = op.source_to_Db(source)
    ldu = int(Db.shape[1])
    kind = op.kind(c)
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        u0.zero_()
        flush.fill_(1)                                    # evict the 126 MB L2 between steps
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _, launches = op.iterate(Db, iters, u0, u1)
        e1.record()
        return e0, e1, launches

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [one_step() for _ in range(args.steps)]
    barrier()
    kernel_ms = [a.elapsed_time(b) for a, b, _ in evs]
    launches = sum(l for _, _, l in evs)
    total_ms = float(np.sum(kernel_ms))

    # ---- e2e: the call a user makes, host buffers in and out ------------------------------------------
    model = gl.ssl.poisson(W, solver="gradient_descent", min_iter=iters, max_iter=iters)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.fit(ti, labels[ti])                              # first fit on this graph: uploads W, builds P/RW on the device
    cold_s = time.perf_counter() - t0
    model.fit(ti, labels[ti])
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        u_host = model.fit(ti, labels[ti])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    h2d = n * c * 8                                        # per fit: the fp64 source term (graph state is cached)
    graph_h2d = (n + 1) * 4 + nnz * 4 + nnz * 8           # once per graph, inside the first fit
    d2h = n * c * 8

    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * iters * args.steps / (total_ms * 1e-3)
    e2e_value = world * iters * e2e_steps / e2e_s
    extras = other_rows(W, labels, ti) if not args.no_extras else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # roofline of the dominant kernel = the persistent iterate: algorithmic bytes per launch / launch duration
    ms_launch = float(np.mean(kernel_ms))
    achieved = algorithmic_bytes(n, nnz, c) * iters / (ms_launch * 1e-3) / 1e9
    # bounded CPU sample of the same workload (oracle port of the reference loop, one core)
    cpu_iters = args.cpu_iters
    cpu_s = min(cpu_reference_loop(W, labels, ti, cpu_iters) for _ in range(2))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 70k nodes, k=10 kNN graph (10 Gaussian blobs), 10 classes, 1 label/class",
                   "n": int(n), "nnz": int(nnz), "classes": c, "iterations_per_step": iters, "ldu": ldu,
                   "kernel": {"dataflow": "poisson_dataflow_kernel", "barrier": "poisson_persistent_kernel",
                              "step": "poisson_step_kernel"}[kind],
                   "gate_every": op.gate(c),
                   "l2": "flushed between steps (256 MiB write); inside a step the 16.8 MB working set is "
                         "L2 resident by construction",
                   "parallelism": "replicas x%d (independent label sets, no collective)" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(kind), "peak_source": peak_src,
                     "traffic_note": "dram__bytes_read+write of one launch from the ncu capture under profiles/ (T=100 "
                                     "iterations per launch there; the traffic is the one-time load of slabs, Db and u - every "
                                     "iteration after the first runs out of L2, which is why achieved > traffic/time)",
                     "bytes_per_iteration": algorithmic_bytes(n, nnz, c)},
        "cpu_baseline": {"value": cpu_iters / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d iterations of the reference loop (ssl.py:667-669, scipy csr_matvecs fp64) on the "
                                   "same graph, best of 2; host has %d cores, scipy SpMM uses 1" % (cpu_iters, os.cpu_count())},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "call": "gl.ssl.poisson(W, solver='gradient_descent').fit -> glb_poisson_graph_fit "
                "(device graph state cached on the gl.graph object after the first fit, as in ssl_trials)",
                "first_fit_value": iters / cold_s, "graph_h2d_bytes_once": int(graph_h2d)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "other_rows": extras,
    }
    assert np.isfinite(u_host).all()
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=1000, help="Poisson iterations per step (one persistent launch)")
    ap.add_argument("--ref-iters", type=int, default=100, help="iterations per step of the CPU reference arm")
    ap.add_argument("--cpu-iters", type=int, default=200, help="bounded CPU sample inside the GPU arm")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other 8(a) rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
