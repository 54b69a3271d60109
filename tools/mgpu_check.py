"""torchrun worker: sharded cgraph + sharded coclust on N GPUs == single-rank result on rank 0's GPU.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metacell_b200 import _lib, synth, dist as mdist      # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    m = synth.make_umi_matrix(5003, 500, n_types=9, median_umis=2000.0, seed=17)         # odd size: ragged last block
    feats = np.arange(0, 500, 2, dtype=np.int32)
    K, kx, kb = 40, 10, 3
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    src, dst, w = mdist.cgraph_sharded(mdist.CudaCGraphEngine(g), K, kx, kb, rank, world)                 # symmetric schedule
    res = mdist.gather_edges(src, dst, w, rank, world, device=local)
    g2 = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
    s2, d2, w2 = mdist.cgraph_sharded(mdist.CudaCGraphEngine(g2), K, kx, kb, rank, world, symmetric=False)   # rectangular schedule
    res2 = mdist.gather_edges(s2, d2, w2, rank, world, device=local)
    ok = True
    if rank == 0:
        g1 = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        g1.knn(K, kx); g1.mutual(kb); g1.prune()
        s1, d1, w1 = g1.edges()
        a, b = set(zip(res[0].tolist(), res[1].tolist())), set(zip(s1.tolist(), d1.tolist()))
        jac = len(a & b) / len(a | b)
        same = res[0].size == s1.size and np.array_equal(res[0], s1) and np.array_equal(res[1], d1) and np.array_equal(res[2], w1)
        print(f"cgraph: sharded {res[0].size} edges, single {s1.size}, jaccard {jac:.6f}, identical {same}", flush=True)
        ok &= jac > 0.9995          # schedules may differ at correlation ties only
        a2 = set(zip(res2[0].tolist(), res2[1].tolist()))
        jac2 = len(a2 & b) / len(a2 | b)
        print(f"cgraph (rectangular schedule): {res2[0].size} edges, jaccard {jac2:.6f}", flush=True)
        ok &= jac2 > 0.9995
    # co-clustering: strided resamples + NCCL reduce
    N = m.ncells
    if rank == 0:
        edges = [torch.from_numpy(x).cuda() for x in (s1.astype(np.int64), d1.astype(np.int64))] + [torch.from_numpy(w1).cuda()]
        n_e = torch.tensor([s1.size], device="cuda")
    else:
        n_e = torch.tensor([0], device="cuda")
    dist.broadcast(n_e, 0)
    if rank != 0:
        edges = [torch.zeros(int(n_e), dtype=torch.int64, device="cuda"), torch.zeros(int(n_e), dtype=torch.int64, device="cuda"),
                 torch.zeros(int(n_e), dtype=torch.float64, device="cuda")]
    for t in edges:
        dist.broadcast(t, 0)
    es, ed, ew = (t.cpu().numpy() for t in edges)
    sets, seeds = synth.resample_sets(N, 0.75, 13, 4)
    cc = _lib.CoClust(ctx, N, es, ed, ew, sets, seeds, 15, 1.05, 10, rank=rank, nranks=world)
    out = mdist.coclust_sharded(cc, rank, world)
    if rank == 0:
        whole = _lib.CoClust(ctx, N, es, ed, ew, sets, seeds, 15, 1.05, 10)
        n1, n2, cnt, ns = whole.pairs()
        same = all(np.array_equal(x, y) for x, y in zip(out, (n1, n2, cnt, ns)))
        print(f"coclust: sharded pairs {out[0].size}, single {n1.size}, identical {same}", flush=True)
        ok &= same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
