#!/usr/bin/env python
"""BASELINE config 5: Poisson iterate on a synthetic 2M-node k=10 graph, row-partitioned over the GPUs of one node.
Launch with torchrun (one rank per GPU) or plain python (one GPU).  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_cfg5.py [--size 2000000]
        --exchange put         halo rows put into the neighbours' label matrices by the step kernel (csrc/slab.cu)   [default]
        --exchange allgather   baseline: step kernel + one NCCL all-gather of the whole label matrix per iteration
        --reorder 0|1          relabel the nodes with the library's RCM locality ordering first                       [1]
    ... --check out.npz   small parity run: partitioned result vs a single-GPU run of the same kernel (used by the tests)
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


def measure(W, rank, world, exchange, reorder, iters, reps, c=10):
    """-> dict (rank 0) with ms per iteration (max over ranks, best of reps) of the chosen exchange."""
    import torch
    import torch.distributed as dist
    from graphlearning_b200 import distributed as gd
    t0 = time.perf_counter()
    if exchange == "put":
        pp = gd.PartitionedPoisson(W, rank=rank, world=world, reorder=bool(reorder), c=c)
        run = lambda T: pp.timed_iterations(T)
    else:
        pp = gd.AllGatherPoisson(W, rank=rank, world=world, reorder=bool(reorder))
        run = lambda T: pp.timed_iterations(c, T)
    t_setup = time.perf_counter() - t0
    run(10)                                                      # warm-up (plan, NCCL communicator, peer mappings)
    best = 1e30
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        ms = run(iters)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        best = min(best, ms)
    n, nnz = W.shape[0], W.nnz
    b_iter = nnz * 8 + (n + 1) * 4 + 3 * n * c * 4
    info = {"exchange": exchange, "reorder": int(reorder), "ms_per_iteration": best / iters, "iterations_per_s": iters / (best * 1e-3),
            "bytes_per_iteration": b_iter, "setup_s": t_setup}
    if exchange == "put":
        sent = torch.tensor([pp.halo_bytes_per_iteration(), pp.part.m, int(pp.part.boundary.sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sent, op=dist.ReduceOp.MAX)
        info.update({"kernel": "slab_step_kernel (halo rows put into peer memory by the kernel)", "put_bytes_per_iteration_per_gpu_max": int(sent[0]),
                     "rows_per_gpu_max": int(sent[1]), "boundary_rows_per_gpu_max": int(sent[2])})
        pp.close()
    else:
        proto = pp._plans[c][1]
        info.update({"kernel": "poisson_step_kernel + ncclAllGather",
                     "allgather_bytes_per_iteration_per_gpu": int((world - 1) * proto.rows_pad * pp._plans[c][2] * 4)})
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=2000000, help="nodes (not --n: torchrun reads that as an abbreviation of its own options)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", default=None)
    ap.add_argument("--reorder", type=int, default=1)
    ap.add_argument("--exchange", default="put", choices=["put", "allgather", "both"])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from graphlearning_b200 import distributed as gd
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    W = build_graph(a.n)
    t_graph = time.perf_counter() - t0
    n, nnz, c = W.shape[0], W.nnz, 10
    if a.check:
        rng = np.random.default_rng(1)
        src = rng.normal(size=(n, c)) * (rng.random((n, 1)) < 0.01)
        pp = gd.PartitionedPoisson(W, rank=rank, world=world, reorder=bool(a.reorder), c=c)
        u = pp.iterate(src, a.iters)
        pp.close()
        ag = gd.AllGatherPoisson(W, rank=rank, world=world, reorder=bool(a.reorder))
        u_ag = ag.iterate(src, a.iters)
        if rank == 0:
            np.savez(a.check, u_put=u, u_allgather=u_ag, src=src)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {"workload": "cfg5: %d-node k=10 graph (uniform points in the unit cube), 10 classes" % n, "n": n, "nnz": int(nnz), "n_gpus": world,
           "iterations": a.iters, "graph_build_s": t_graph, "runs": []}
    for ex in (["put", "allgather"] if a.exchange == "both" else [a.exchange]):
        info = measure(W, rank, world, ex, a.reorder, a.iters, a.reps, c)
        info["achieved_GBs_all_gpus"] = info["bytes_per_iteration"] * info["iterations_per_s"] / 1e9
        info["frac_of_hbm_peak_x_gpus"] = info["achieved_GBs_all_gpus"] / (peak * world)
        out["runs"].append(info)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
