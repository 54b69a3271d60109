"""C4-shaped scale run (BASELINE.json configs[3]): N cells x 1,000 feature genes, K = 150, one process per GPU.  Every rank
generates and uploads ITS OWN cell range (sharded ingest) and calls the rank-collective build of libmcb200.so
(mcb_cgraph_build_sharded: NCCL inside the library); rank 0 prints timings and size-independent checks.
usage: torchrun ... tools/c4_scale.py [ncells] [K]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metacell_b200 import _lib, synth      # noqa: E402


def main():
    ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(_lib.Comm.unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    comm = _lib.Comm(ctx, world, rank, idt.cpu().numpy().tobytes())
    per = ncells // world
    t0 = time.time()
    m = synth.make_umi_matrix(per, 1000, n_types=40, median_umis=1500.0, seed=1234, cell_seed=rank)
    t_gen = time.time() - t0
    x = m.data.astype(np.float64)
    N = per * world
    feats = np.arange(1000, dtype=np.int32)
    local_csc = _lib.Csc.upload(ctx, 1000, m.indptr, m.indices, x)            # this rank's cells, resident
    times, times_e2e = [], []
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        g = comm.build_sharded(local_csc, rank * per, feats, K, 10, 3, 7.0)
        ctx.sync(); dist.barrier()
        times.append(time.time() - t0)
        if it == 0:
            stats = g.knn_stats()
            col, cor = g.download_knn()
            src, dst, w = g.edges()
            M, row0, n_edges = g.M, g.row0, g.n_edges
        g.destroy()
    for it in range(2):                 # host in -> device-resident sharded edge table (upload of the shard inside)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        c2 = _lib.Csc.upload(ctx, 1000, m.indptr, m.indices, x)
        g = comm.build_sharded(c2, rank * per, feats, K, 10, 3, 7.0)
        ctx.sync(); dist.barrier()
        times_e2e.append(time.time() - t0)
        g.destroy(); c2.free()
    # size-independent checks on this rank's rows
    ok = True
    ok &= bool(np.all(col[:, 0] == np.arange(row0, row0 + col.shape[0])))                  # self at rank 1
    ok &= bool(np.all(np.diff(cor[:, 1:], axis=1) <= 0))                                   # ranked by correlation
    ok &= bool(np.all([len(set(r.tolist())) == r.size for r in col[:: max(1, col.shape[0] // 64)]]))   # no duplicate neighbours
    deg = np.bincount(src - src.min(), minlength=1)
    ok &= int(deg.max()) <= K and bool(np.all((w > 0) & (w <= 1.0))) and bool(np.all(src != dst))
    tot_edges = torch.tensor([n_edges], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot_edges)
    indeg = torch.zeros(N, dtype=torch.int64, device="cuda")
    indeg.index_add_(0, torch.from_numpy(dst.astype(np.int64)).cuda(), torch.ones(dst.size, dtype=torch.int64, device="cuda"))
    dist.all_reduce(indeg)
    ok &= int(indeg.max()) <= 3 * K                                                        # in-degree cut across ranks
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    free, total = torch.cuda.mem_get_info()
    if rank == 0:
        print(json.dumps({"config": f"C4-shaped: {N} cells x 1000 feature genes, K={K}, {world} x B200, one process per GPU, NCCL inside the library",
                          "gen_s_per_rank": round(t_gen, 1), "build_s": [round(t, 3) for t in times], "cells_per_s": N / min(times[1:]),
                          "host_in_s": [round(t, 3) for t in times_e2e], "cells_per_s_host_in": N / min(times_e2e),
                          "valid_cells": M, "edges": int(tot_edges), "edges_per_cell": int(tot_edges) / N, "max_in_degree": int(indeg.max()),
                          "knn_filter": stats, "device_GiB_in_use": round((total - free) / 2**30, 1), "checks": "ok" if int(okt) else "FAILED"}), flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(okt) else 1)


if __name__ == "__main__":
    main()
