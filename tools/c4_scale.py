"""C4-shaped scale run (BASELINE.json configs[3]): N cells x 1,000 feature genes, K = 150, rows and the symmetric tile
list sharded over all ranks.  Every rank generates its own shard of cells, the CSC shards are all-gathered over NCCL,
then the balanced graph is built; rank 0 prints timings and size-independent checks.
usage: torchrun ... tools/c4_scale.py [ncells] [K]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metacell_b200 import _lib, synth, dist as mdist      # noqa: E402


def main():
    ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    per = ncells // world
    t0 = time.time()
    m = synth.make_umi_matrix(per, 1000, n_types=40, median_umis=1500.0, seed=1234, cell_seed=rank)
    t_gen = time.time() - t0
    # all-gather the shards into one device-resident CSC per rank
    nnz = torch.tensor([m.nnz], dtype=torch.int64, device="cuda")
    all_nnz = [torch.zeros_like(nnz) for _ in range(world)]
    dist.all_gather(all_nnz, nnz)
    all_nnz = [int(x) for x in all_nnz]
    mx = max(all_nnz)
    def gather(a, dtype):
        t = torch.zeros(mx, dtype=dtype, device="cuda")
        t[: a.size] = torch.from_numpy(a).cuda().to(dtype)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat([outs[r][: all_nnz[r]] for r in range(world)])
    d_i = gather(m.indices.astype(np.int32), torch.int32)
    d_x = gather(m.data.astype(np.int32), torch.int32)
    lens = torch.from_numpy(np.diff(m.indptr)).cuda()
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    d_p = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(torch.cat(all_lens), 0)])
    N = per * world
    csc = _lib.Csc.wrap_device(ctx, 1000, N, int(d_p[-1]), d_p.data_ptr(), d_i.data_ptr(), d_x.data_ptr(), keep=(d_p, d_i, d_x))
    feats = np.arange(1000, dtype=np.int32)
    times = []
    for it in range(2):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        eng = mdist.CudaCGraphEngine(g)
        n_edges = mdist.cgraph_sharded(eng, K, 10, 3, rank, world, fetch_edges=False)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.time() - t0)
        if it == 0:
            stats = g.knn_stats()
            col, cor = g.download_knn()
            src, dst, w = g.edges()
            M, row0 = g.M, g.row0
        g.destroy()
    # size-independent checks on this rank's rows
    ok = True
    ok &= bool(np.all(col[:, 0] == np.arange(row0, row0 + col.shape[0])))                  # self at rank 1
    ok &= bool(np.all(np.diff(cor[:, 1:], axis=1) <= 0))                                   # ranked by correlation
    ok &= bool(np.all([len(set(r.tolist())) == r.size for r in col[:: max(1, col.shape[0] // 64)]]))   # no duplicate neighbours
    deg = np.bincount(src - src.min(), minlength=1)
    ok &= int(deg.max()) <= K and bool(np.all((w > 0) & (w <= 1.0)))
    tot_edges = torch.tensor([n_edges], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot_edges)
    okt = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    mem = torch.cuda.max_memory_allocated() / 2**30
    free, total = torch.cuda.mem_get_info()
    if rank == 0:
        print(f"C4-shaped: N={N} cells, K={K}, world={world}: gen {t_gen:.1f}s/rank, build {times[0]:.2f}s (first) {times[1]:.2f}s (steady) "
              f"-> {N / times[1]:.0f} cells/s; valid cells {M}, edges {int(tot_edges)} ({int(tot_edges) / N:.1f}/cell), "
              f"knn filter {stats}, device memory in use {(total - free) / 2**30:.1f} GiB, checks {'ok' if int(okt) else 'FAILED'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(okt) else 1)


if __name__ == "__main__":
    main()
