#!/usr/bin/env python
"""bench.py - Poisson SpMM iterations/sec on the 70k-node k=10 graph with 10 classes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--iters I]

One "step" = I iterations (default 1000) of u <- Db + P u on the device-resident graph = ONE launch of the
persistent kernel.  `value` is whole-job iterations/s with all inputs resident in HBM; `e2e` is the same
metric through the reference-facing API gl.ssl.poisson(...).fit(train_ind, train_labels) with HOST buffers: every
fit uploads the labelled rows of the source term (it is zero elsewhere, ssl.py:619-622) and train_ind, and reads the
n x c fp64 scores back into host memory (a page-locked array of the library's pool) inside the timed region; the
graph itself is uploaded and normalised once, in the first fit (`first_fit_ms`).  N>1: the 70k graph fits one GPU, so
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
    (~10 s, untimed setup, identical for both arms).  The literal config-2 graph (128-d features, built by the GPU kNN
    search) is measured too and reported as `cfg2_literal_128d`: isotropic 128-d Gaussians give hub nodes with > 1000
    neighbours, which real kNN graphs (MNIST: max 45) do not have."""
    from oracle import gl_oracle as orc
    from scipy import sparse
    X, labels = orc.synthetic_blobs(N_NODES, 8, c=N_CLASSES, seed=seed)
    ind, dist = orc.knnsearch(X.astype(np.float64), K_NN + 1, method="kdtree")
    W = orc.knn_weights(ind, dist, K_NN)
    return sparse.csr_matrix(W), labels


WORKLOAD = "cfg2 (synthetic, d=8): 70k points of 10 Gaussian blobs in R^8, k=10 kNN graph, 10 classes, 1 label/class"

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, ncu --set full
# (profiles/r2_dataflow_T1000_ncu_details.txt: 19.34 MB + 0.28 MB; profiles/r2_dataflow_pipe_gate1_ncu_details.txt, another
# launch length: 19.67 MB + 0.35 MB - the traffic does not depend on the number of iterations of the launch)
NCU_DRAM_BYTES_PER_LAUNCH = {"dataflow": 19336704 + 284672, "barrier": 17295360 + 296704, "step": None}


def timed(fn, reps=1):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def literal_cfg2(labels_unused=None):
    """The literal config 2: 70 000 x 128 features -> GPU kNN search -> Poisson iterate on the graph it builds."""
    import torch
    from graphlearning_b200 import device as gdev, knn_gpu
    from oracle import gl_oracle as orc
    X, lab = orc.synthetic_blobs(N_NODES, 128, c=N_CLASSES, seed=0)
    X = X.astype(np.float64)
    knn_gpu.knnsearch_gpu(X[:4096], K_NN + 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ind, dist_ = knn_gpu.knnsearch_gpu(X, K_NN + 1)
    torch.cuda.synchronize(); t_knn = time.perf_counter() - t0
    fallback = knn_gpu.last_stats.get("fallback_rows")
    Wh = orc.knn_weights(ind, dist_, K_NN)
    deg = np.diff(Wh.indptr)
    th = orc.one_per_class(lab, rate=1, seed=0)
    op = gdev.PoissonOperator(Wh, reorder=True)
    Db = op.source_to_Db(orc.poisson_source(N_NODES, th, lab[th])[0])
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    best = 1e30
    for _ in range(3):
        u0.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); op.iterate(Db, 500, u0, u1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    b = algorithmic_bytes(N_NODES, Wh.nnz, N_CLASSES)
    # CPU baseline of the search, measured here: the reference's exact branch (cKDTree, weightmatrix.py:349-352) on a
    # sub-sample of query rows against all points, extrapolated to n rows
    from scipy import spatial
    rows = 100
    tree = spatial.cKDTree(X)
    t_tree, _ = timed(lambda: tree.query(X[:rows], k=K_NN + 1))
    return {
        "knn_70k_x_128_k10": {"seconds_host_to_host": t_knn, "fp32_TFLOPs_on_2n2d": 2.0 * N_NODES * N_NODES * 128 / t_knn / 1e12,
                              "fallback_rows": fallback,
                              "reference_cpu": {"seconds_extrapolated": t_tree / rows * N_NODES, "measured": "%d query rows against all 70 000 points in %.2f s, "
                                                "scipy cKDTree, 1 thread (the reference's call, weightmatrix.py:351-352), scaled to 70 000 rows" % (rows, t_tree)}},
        "poisson_iterate": {"nnz": int(Wh.nnz), "max_row": int(deg.max()), "p99_row": int(np.percentile(deg, 99)), "kernel": op.kind(N_CLASSES),
                            "gate_every": op.gate(N_CLASSES), "us_per_iteration": best * 1e3 / 500, "iterations_per_s": 500 / (best * 1e-3),
                            "bytes_per_iteration": b, "achieved_GBs": b * 500 / (best * 1e-3) / 1e9},
    }


def reference_use_cuda(W, labels, ti, iters):
    """The reference's OWN GPU path (`use_cuda=True`: torch COO fp32 sparse.addmm, ssl.py:649-663) on this B200 - the
    "GPU baseline to beat" of SURVEY 8a row a14 / BASELINE.md 4.3 - restated in oracle/gl_oracle.py."""
    import torch
    from oracle import gl_oracle as orc
    orc.poisson_gd_use_cuda(W, ti, labels[ti], min_iter=5, max_iter=5)                       # warm-up (cuSPARSE handles)
    torch.cuda.synchronize()
    t_host, _ = timed(lambda: (orc.poisson_gd_use_cuda(W, ti, labels[ti], min_iter=iters, max_iter=iters), torch.cuda.synchronize()))
    t_dev, _ = timed(lambda: (orc.poisson_gd_use_cuda(W, ti, labels[ti], min_iter=iters, max_iter=iters, host_mixing=False), torch.cuda.synchronize()))
    t_setup, _ = timed(lambda: (orc.poisson_gd_use_cuda(W, ti, labels[ti], min_iter=0, max_iter=0), torch.cuda.synchronize()))
    return {"iterations": iters,
            "iterations_per_s_as_shipped": iters / max(t_host - t_setup, 1e-9),
            "iterations_per_s_addmm_only": iters / max(t_dev - t_setup, 1e-9),
            "note": "as shipped: torch.sparse.addmm on the device + v = RW*v and np.max on the host every iteration (ssl.py:657-660); "
                    "addmm only: the same loop without the host-side stopping vector.  Setup (P to COO, uploads) subtracted: %.3f s" % t_setup}


def other_rows(W, labels, ti):
    """Short measurements of the other rows of SURVEY.md 8(a) on this box (rank 0, untimed part of the bench line), each
    next to the reference's CPU implementation of the same step measured here (bounded samples)."""
    import torch
    import graphlearning_b200 as gl
    from oracle import gl_oracle as orc, c_oracle
    out = {}
    try:
        m = gl.ssl.laplace(W)
        t5 = orc.one_per_class(labels, rate=5, seed=0)
        m.fit(t5, labels[t5])
        from graphlearning_b200 import device as _gdev
        _gdev.pinned.wait()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(t5, labels[t5])
        t = time.perf_counter() - t0
        out["laplace_cg_fit"] = {"seconds": t, "cg_iterations": int(m.iterations), "gpu_launches": int(m.gpu_launches),
                                 "note": "gl.ssl.laplace(W).fit, 5 labels/class, tol 1e-5, host buffers in and out"}
        out["laplace_cg_fit"]["cg_device_ms"] = m.cg_info["device_ms"]
        # config 3: 60 000 x 512 features, k = 20, Laplace learning (CG, tol 1e-5, 5 labels/class): GPU kNN build, device
        # assembly of the Dirichlet system, CG.  B_cg = SURVEY 8(d): nnz*8 + (n+1)*4 + 11*n*c*4 (its fp32 accounting; the
        # solver stores fp64 for parity, so it moves about twice that)
        X3, lab3 = orc.synthetic_blobs(60000, 512, c=N_CLASSES, seed=0)
        t0 = time.perf_counter(); W3 = gl.weightmatrix.knn(X3.astype(np.float64), 20); t_graph = time.perf_counter() - t0
        t3 = orc.one_per_class(lab3, rate=5, seed=0)
        m3 = gl.ssl.laplace(W3)
        m3.fit(t3, lab3[t3])
        _gdev.pinned.wait()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m3.fit(t3, lab3[t3])
        t_fit = time.perf_counter() - t0
        ci = m3.cg_info
        b_cg = ci["system_nnz"] * 8 + (ci["unknowns"] + 1) * 4 + 11 * ci["unknowns"] * N_CLASSES * 4
        us_it = 1e3 * ci["device_ms"] / max(1, m3.iterations)
        out["cfg3_laplace_60k_x_512_k20"] = {
            "graph_build_seconds": t_graph, "nnz": int(W3.nnz), "fit_seconds_host_to_host": t_fit, "cg_iterations": int(m3.iterations),
            "cg_device_ms": ci["device_ms"], "us_per_cg_iteration": us_it, "B_cg_bytes": b_cg,
            "achieved_GBs_on_B_cg": b_cg / (us_it * 1e-6) / 1e9, "frac_of_hbm_peak": b_cg / (us_it * 1e-6) / 1e9 / hbm_peak()[0],
            "accuracy_percent": float(gl.ssl.ssl_accuracy(m3.predict(), lab3, t3)), "gpu_launches": int(m3.gpu_launches)}
        # config 4: 50 eigenpairs of the normalised Laplacian (graph.eigen_decomp on the block kernels of spectral.cu)
        G = gl.graph(W)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vals, vecs = G.eigen_decomp(normalization="normalized", k=50)
        t = time.perf_counter() - t0
        t_arpack, ref = timed(lambda: orc.eigen_decomp(W, normalization="normalized", k=50))
        out["eigen_decomp_k50"] = {"seconds": t, "spmm_launches": int(G.eigen_info["spmm"]), "residual": float(G.eigen_info["residual"]),
                                   "lambda_50": float(vals[-1]), "max_abs_eigenvalue_diff_vs_reference": float(np.max(np.abs(vals - ref[0]))),
                                   "reference_cpu_seconds": t_arpack, "reference_cpu": "scipy ARPACK svds as graph.py:728-746, measured here"}
        # p-Laplace / AMLE sweeps (bit-exact Gauss-Seidel / Jacobi of c_code/lp_iterate.cpp), one class against the rest
        val = (labels[t5] == 0).astype(np.float64)
        G.plaplace(t5, val, 3, max_num_it=30)
        for name, fn in (("plaplace_p3_fast", lambda: G.plaplace(t5, val, 3)), ("amle_weighted", lambda: G.amle(t5, val, tol=1e-3, max_num_it=300))):
            t0 = time.perf_counter(); fn(); t = time.perf_counter() - t0
            out[name] = {"seconds_host_to_host": t, "sweeps": int(G.sweeps), "us_per_sweep": 1e6 * t / max(1, G.sweeps)}
        I, J, V = orc.ccode_triplets(W)
        sweeps = 40
        t_c, _ = timed(lambda: c_oracle.lip_iterate(np.zeros(len(labels)), J, I, V, t5.astype(np.int32), val, sweeps, 0.0, 1.0 / 3, 2.0 / 3))
        out["plaplace_p3_fast"]["reference_cpu_us_per_sweep"] = 1e6 * t_c / sweeps
        out["plaplace_p3_fast"]["reference_cpu"] = "lip_iterate_main of c_code/lp_iterate.cpp (plain-C restatement, gcc -O2), %d sweeps measured here, 1 core" % sweeps
        mp = gl.ssl.plaplace(W, p=3)                                # one-vs-rest over the 10 classes, batched sweep kernel
        t0 = time.perf_counter(); mp.fit(t5, labels[t5]); t = time.perf_counter() - t0
        out["ssl_plaplace_p3_fit_10_classes"] = {"seconds": t, "sweeps_per_class": [int(x) for x in mp.graph.sweeps]}
        # wider label matrices: B independent label sets (the trials of ssl_trials) batched as 10 B columns of ONE iterate.  A
        # gather costs one L1 wavefront per label row whatever its width up to 128 bytes, so the cost per label set drops
        from graphlearning_b200 import device as gdev
        src1 = orc.poisson_source(len(labels), ti, labels[ti])[0]
        opw = gdev.PoissonOperator(W, reorder=True)
        wide = {}
        for B in (1, 2, 4):
            cB = N_CLASSES * B
            Dbw = opw.source_to_Db(np.tile(src1, (1, B)))
            a0 = torch.zeros_like(Dbw); a1 = torch.zeros_like(Dbw)
            opw.iterate(Dbw, 50, a0, a1, c=cB)
            best = 1e30
            for _ in range(3):
                a0.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); opw.iterate(Dbw, 500, a0, a1, c=cB); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            bB = algorithmic_bytes(len(labels), W.nnz, cB)
            wide["label_sets_%d" % B] = {"columns": cB, "kernel": opw.kind(cB), "us_per_iteration": best * 1e3 / 500,
                                         "label_set_iterations_per_s": B * 500 / (best * 1e-3), "bytes_per_iteration": bB,
                                         "frac_of_hbm_peak": bB * 500 / (best * 1e-3) / 1e9 / hbm_peak()[0]}
        out["batched_label_sets"] = wide
        del opw, Dbw, a0, a1
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
        "config": {"workload": WORKLOAD, "n": int(W.shape[0]), "nnz": int(W.nnz), "classes": N_CLASSES, "iterations_per_step": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def parity_gate(W, labels, ti):
    """every number ships with its parity check on the benched graph - scores after T = 50 iterations
    against the fp64 oracle (max|u - u_ref| / max|u_ref| <= 1e-5) and identical predict() labels."""
    import graphlearning_b200 as gl
    from oracle import gl_oracle as orc, c_oracle
    T = 50
    model = gl.ssl.poisson(W, solver="gradient_descent", min_iter=T, max_iter=T)
    u = model.fit(ti, labels[ti])
    s = orc.poisson_gd_setup(W, ti, labels[ti])
    u_ref = c_oracle.poisson_iterate(s["P"], np.asarray(s["Db"]), T)
    err = float(np.max(np.abs(u - u_ref)) / np.max(np.abs(u_ref)))
    same = bool(np.array_equal(model.predict(), orc.predict(u_ref)))
    # default stopping rule: the iteration count must be the reference's (mixing of v <- RW v in fp64, ssl.py:667,669)
    T_ref = 0
    v, vinf, RW = s["v"], s["vinf"], s["RW"]
    while (T_ref < 50 or np.max(np.absolute(v - vinf)) > 1 / s["n"]) and T_ref < 1000:
        v = RW * v; T_ref += 1
    return {"T": T, "max_rel_err_vs_fp64_oracle": err, "tolerance": 1e-5, "predict_equal": same, "T_default_rule_reference": T_ref,
            "ok": bool(err <= 1e-5 and same)}


def cfg5_rowpart(rank, world, full):
    """BASELINE config 5 through graphlearning_b200.distributed: the 2M-node graph row-partitioned over the ranks, halo
    rows put into the neighbours' label matrices by the step kernel (csrc/slab.cu); the all-gather variant of the north
    star is measured next to it.  N = 1 runs the same kernel on one slab (no peers), so the driver's 1/2/4/8 curve is
    this row.  Returns a dict on rank 0."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_cfg5
    from graphlearning_b200 import distributed as gd
    n = 2000000
    t0 = time.perf_counter()
    W = bench_cfg5.build_graph(n)
    t_graph = time.perf_counter() - t0
    peak = hbm_peak()[0]
    out = {"workload": "cfg5: 2M points uniform in the unit cube, k=10 kNN graph, 10 classes, T=100 iterations per run",
           "n": n, "nnz": int(W.nnz), "n_gpus": world, "graph_build_s": t_graph}
    put = bench_cfg5.measure(W, rank, world, "put", 1, 100, 5)
    out.update({"iterations_per_s": put["iterations_per_s"], "ms_per_iteration": put["ms_per_iteration"],
                "achieved_GBs_all_gpus": put["bytes_per_iteration"] * put["iterations_per_s"] / 1e9,
                "frac_of_hbm_peak_per_gpu": put["bytes_per_iteration"] * put["iterations_per_s"] / 1e9 / (peak * world),
                "exchange": "halo rows put into peer memory (NVLink, CUDA IPC) by slab_step_kernel, interior rows overlap",
                "put_bytes_per_iteration_per_gpu_max": put.get("put_bytes_per_iteration_per_gpu_max"),
                "boundary_rows_per_gpu_max": put.get("boundary_rows_per_gpu_max"), "rows_per_gpu_max": put.get("rows_per_gpu_max"),
                "setup_s": put["setup_s"]})
    if world > 1 or full:
        ag = bench_cfg5.measure(W, rank, world, "allgather", 1, 100, 2)
        out["allgather_baseline"] = {"iterations_per_s": ag["iterations_per_s"], "kernel": ag["kernel"],
                                     "allgather_bytes_per_iteration_per_gpu": ag.get("allgather_bytes_per_iteration_per_gpu")}
    if world > 1:
        # parity: the partitioned run against ONE slab holding the whole (smaller) graph, bitwise
        Ws = bench_cfg5.build_graph(100000)
        src = np.random.default_rng(1).normal(size=(Ws.shape[0], N_CLASSES)) * (np.random.default_rng(2).random((Ws.shape[0], 1)) < 0.01)
        pp = gd.PartitionedPoisson(Ws, rank=rank, world=world, reorder=True, c=N_CLASSES)
        u_part = pp.iterate(src, 12)
        pp.close()
        if rank == 0:
            one = gd.PartitionedPoisson(Ws, rank=0, world=1, reorder=True, c=N_CLASSES)
            u_one = one.iterate(src, 12)
            one.close()
            out["parity_vs_single_gpu"] = bool(np.array_equal(u_part, u_one))
        dist.barrier()
    return out


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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

    op = gdev.PoissonOperator(W, kind=os.environ.get("GLB_BENCH_KIND", "auto"), reorder=True)   # profiler runs pin the kernel: the plan's trial timing is meaningless under ncu
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
    model.fit(ti, labels[ti])                              # first fit on this graph: uploads W, builds P/RW, ordering and plan on the device
    cold_s = time.perf_counter() - t0
    model.fit(ti, labels[ti])
    from graphlearning_b200 import device as _gdev
    _gdev.pinned.wait()                                    # steady state: the pool's background thread has pinned the result buffers
    model.fit(ti, labels[ti])
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        u_host = model.fit(ti, labels[ti])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches += e2e_steps * int(model.gpu_launches)
    # first fit on ANOTHER graph object in the now warm process: what a new graph costs (upload, transposition, ordering, slab
    # build, plan) without the one-time costs of the process (CUDA module loading, first pinned allocation)
    W2 = W.copy()
    model2 = gl.ssl.poisson(W2, solver="gradient_descent", min_iter=iters, max_iter=iters)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model2.fit(ti, labels[ti])
    cold2_s = time.perf_counter() - t0
    launches += int(model2.gpu_launches)
    del model2, W2
    # the same call with the reference's defaults (min_iter=50, max_iter=1000: the stopping rule runs on the device too)
    dmodel = gl.ssl.poisson(W, solver="gradient_descent")
    dmodel.fit(ti, labels[ti])
    _gdev.pinned.wait()
    dmodel.fit(ti, labels[ti])
    barrier()
    default_ms = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        dmodel.fit(ti, labels[ti])                         # synchronous: the scores are in host memory when it returns
        default_ms.append(1e3 * (time.perf_counter() - t0))
    e2e_default_s = 1e-3 * float(np.mean(default_ms))
    T_default = int(dmodel.iterations)
    launches += e2e_steps * int(dmodel.gpu_launches)
    clocks = sampler.stop() if rank == 0 else None
    h2d = len(ti) * (8 + c * 8) + len(ti) * 8              # per fit: the labelled rows of the fp64 source term + their indices, train_ind (graph state is cached)
    graph_h2d = (n + 1) * 4 + nnz * 4 + nnz * 8           # once per graph, inside the first fit
    d2h = n * c * 8

    if world > 1:
        t = torch.tensor([total_ms, e2e_s, e2e_default_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_default_s = float(t[0]), float(t[1]), float(t[2])
        dist.barrier()
    del flush
    torch.cuda.empty_cache()
    rowpart = None
    if not args.no_cfg5:
        try:
            rowpart = cfg5_rowpart(rank, world, full=not args.no_extras)
        except Exception as e:
            rowpart = {"error": repr(e)}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    value = world * iters * args.steps / (total_ms * 1e-3)
    e2e_value = world * iters * e2e_steps / e2e_s
    parity = parity_gate(W, labels, ti)
    extras = None
    if not args.no_extras and world == 1:
        extras = other_rows(W, labels, ti)
        try:
            extras["cfg2_literal_128d"] = literal_cfg2()
            extras["reference_gpu_path_use_cuda"] = reference_use_cuda(W, labels, ti, 200)
        except Exception as e:
            extras["error_2"] = repr(e)
    peak, peak_src = hbm_peak()
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
        "config": {"workload": WORKLOAD,
                   "n": int(n), "nnz": int(nnz), "classes": c, "iterations_per_step": iters, "ldu": ldu,
                   "kernel": {"dataflow": "poisson_dataflow_kernel", "barrier": "poisson_persistent_kernel",
                              "step": "poisson_step_kernel"}[kind],
                   "gate_every": op.gate(c), "node_ordering": "reverse Cuthill-McKee (library, per graph)",
                   "l2": "flushed between steps (256 MiB write); inside a step the 16.8 MB working set is "
                         "L2 resident by construction",
                   "parallelism": "replicas x%d (independent label sets, no collective); the row-partitioned multi-GPU "
                                  "iterate of config 5 is `cfg5_rowpart`" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(kind), "peak_source": peak_src,
                     "traffic_note": "dram__bytes_read+write of ONE launch (all iterations of a step) from the ncu --set full "
                                     "capture under profiles/: the one-time load of slabs, Db and u - every iteration after the "
                                     "first runs out of L2, so the DRAM traffic of a launch does not grow with its iteration count "
                                     "(algorithmic bytes per launch = bytes_per_iteration x iterations_per_step)",
                     "bytes_per_iteration": algorithmic_bytes(n, nnz, c)},
        "cpu_baseline": {"value": cpu_iters / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d iterations of the reference loop (ssl.py:667-669, scipy csr_matvecs fp64) on the "
                                   "same graph, best of 2; host has %d cores, scipy SpMM uses 1" % (cpu_iters, os.cpu_count())},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "call": "gl.ssl.poisson(W, solver='gradient_descent', min_iter=max_iter=%d).fit -> glb_poisson_graph_fit_rows "
                "(host buffers: labels in, the n x c fp64 scores out into a page-locked numpy array; the source term is zero outside "
                "the labelled rows, so only those rows are uploaded; device graph state cached on the gl.graph object after the "
                "first fit, as in ssl_trials)" % iters,
                "first_fit_value": iters / cold_s, "first_fit_ms": 1e3 * cold_s, "first_fit_ms_next_graph": 1e3 * cold2_s, "graph_h2d_bytes_once": int(graph_h2d)},
        "e2e_default": {"call": "gl.ssl.poisson(W, solver='gradient_descent').fit with the reference's defaults min_iter=50, max_iter=1000: "
                                "the stopping vector v <- RW v (fp64) is iterated on the device as well", "T": T_default,
                        "ms_per_fit": 1e3 * e2e_default_s, "ms_each_fit": [round(x, 3) for x in default_ms],
                        "value": T_default / e2e_default_s, "unit": UNIT},
        "parity": parity,
        "cfg5_rowpart": rowpart,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "other_rows": extras,
    }
    assert np.isfinite(u_host).all()
    if not parity["ok"]:
        emit(out)
        raise SystemExit("parity gate failed: %r" % (parity,))
    emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def emit(obj):
    """The ONE JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    # stdout carries exactly one line (rank 0's JSON): whatever libraries print there (NCCL's "NCCL version ..." banner under
    # torchrun) is sent to stderr by pointing file descriptor 1 at 2 for the duration of the run
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=1000, help="Poisson iterations per step (one persistent launch)")
    ap.add_argument("--ref-iters", type=int, default=100, help="iterations per step of the CPU reference arm")
    ap.add_argument("--cpu-iters", type=int, default=200, help="bounded CPU sample inside the GPU arm")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other 8(a) rows")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the row-partitioned config-5 measurement")
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
