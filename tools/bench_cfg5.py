#!/usr/bin/env python
"""BASELINE config 5: Poisson iterate on a synthetic 2M-node k=10 graph, row-partitioned over the GPUs of one node
with one all-gather of the label matrix per iteration.  Launch with torchrun (one rank per GPU) or plain python
(one GPU).  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cfg5.py [--size 2000000]
    ... --check out.npz   small parity run: partitioned result vs the single-GPU step kernel (used by the tests)
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_graph(n, k=10, seed=0):
    """Points uniform in [0,1]^3, exact kNN by cKDTree (3-d is cheap), gaussian weights, symmetrised (SURVEY 8d cfg 5)."""
    from scipy import spatial
    from oracle import gl_oracle as orc
    X = np.random.default_rng(seed).random((n, 3))
    dist, ind = spatial.cKDTree(X).query(X, k=k + 1, workers=-1)
    return orc.knn_weights(ind, dist, k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=2000000, help="nodes (not --n: torchrun reads that as an abbreviation of its own options)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", default=None)
    ap.add_argument("--reorder", type=int, default=0, help="1: relabel the nodes with the library's RCM locality ordering first")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from graphlearning_b200 import distributed as gd, device as gdev
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    W = build_graph(a.n)
    t_graph = time.perf_counter() - t0
    t_order = 0.0
    if a.reorder:
        import ctypes
        from graphlearning_b200 import _lib
        t0 = time.perf_counter()
        rp = np.ascontiguousarray(W.indptr, dtype=np.int32); ci = np.ascontiguousarray(W.indices, dtype=np.int32)
        perm = np.empty(W.shape[0], dtype=np.int32)
        _lib.call("glb_locality_order_host", ctypes.c_void_p(rp.ctypes.data), ctypes.c_void_p(ci.ctypes.data), W.shape[0],
                  ctypes.c_void_p(perm.ctypes.data))
        W = W[perm][:, perm].tocsr()
        t_order = time.perf_counter() - t0
    n, nnz, c = W.shape[0], W.nnz, 10
    pp = gd.PartitionedPoisson(W, rank=rank, world=world)
    if a.check:
        rng = np.random.default_rng(1)
        src = rng.normal(size=(n, c)) * (rng.random((n, 1)) < 0.01)
        u = pp.iterate(src, a.iters)
        if rank == 0:
            op = gdev.PoissonOperator(W, kind="step")
            ref = op.unpack(op.iterate(op.source_to_Db(src), a.iters)[0], c).cpu().numpy()
            np.savez(a.check, u_partitioned=u, u_single=ref)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    pp.timed_iterations(c, 10)                                   # warm-up (plan, NCCL communicator)
    best = 1e30
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        ms = pp.timed_iterations(c, a.iters)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        best = min(best, ms)
    if rank == 0:
        b_iter = nnz * 8 + (n + 1) * 4 + 3 * n * c * 4
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        its = a.iters / (best * 1e-3)
        print(json.dumps({"workload": "cfg5: %d-node k=10 graph (uniform points in the unit cube), 10 classes" % n, "n": n, "nnz": int(nnz),
                          "n_gpus": world, "iterations": a.iters, "ms_per_iteration": best / a.iters, "iterations_per_s": its,
                          "bytes_per_iteration": b_iter, "achieved_GBs_all_gpus": b_iter * its / 1e9,
                          "frac_of_hbm_peak_x_gpus": b_iter * its / 1e9 / (peak * world),
                          "allgather_bytes_per_iteration_per_gpu": int((world - 1) * pp._plans[c][1].rows_pad * pp._plans[c][2] * 4),
                          "graph_build_s": t_graph, "reorder": a.reorder, "reorder_s": t_order, "kernel": "poisson_step_kernel + ncclAllGather"}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
