"""In-library multi-GPU checks (NCCL inside libmcb200.so; no torch on the data path).

  python tools/multi_check.py                  one process driving every visible GPU (mcb_multi): mcb_cgraph_bknn_multi and
                                               mcb_coclust_multi against the single-device entry points
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_check.py
                                               one process per GPU (mcb_comm): mcb_cgraph_build_sharded +
                                               mcb_cgraph_gather_edges + mcb_coclust_reduce; the launcher (gloo) only
                                               carries the 128-byte NCCL id and the pass/fail flag
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metacell_b200 import _lib, synth      # noqa: E402

K, KX, KB = 40, 10, 3


def inputs():
    m = synth.make_umi_matrix(5003, 700, n_types=9, median_umis=2000.0, seed=17)         # odd size: ragged last block
    feats = np.sort(np.random.default_rng(3).choice(700, 250, replace=False)).astype(np.int32)
    data = m.data.copy()
    for c in (11, 2500, 5002):                                                            # degenerate cells on different ranks
        sl = slice(m.indptr[c], m.indptr[c + 1])
        data[sl] = 0
    return m, feats, data.astype(np.float64)


def jaccard(a, b):
    sa, sb = set(zip(a[0].tolist(), a[1].tolist())), set(zip(b[0].tolist(), b[1].tolist()))
    return len(sa & sb) / max(1, len(sa | sb))


def single_process():
    import torch
    n = torch.cuda.device_count()
    m, feats, x = inputs()
    ctx = _lib.Context(0)
    ok = True
    for dsamp in (False, True):
        ref = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, x, feats, K, KX, KB, 7.0, dsamp=dsamp, seed=5)
        for ndev in sorted({1, 2, n} & set(range(1, n + 1))):
            mg = _lib.Multi(list(range(ndev)))
            got = _lib.cgraph_bknn_oneshot(mg, m.ngenes, m.indptr, m.indices, x, feats, K, KX, KB, 7.0, dsamp=dsamp, seed=5)
            same = all(np.array_equal(a, b) for a, b in zip(got[:3], ref[:3])) and got[3] == ref[3]
            jac = jaccard(got, ref)
            print(f"multi bknn dsamp={dsamp} devices={ndev}: edges {got[0].size} vs {ref[0].size}, valid {got[3]}, jaccard {jac:.6f}, identical {same}", flush=True)
            ok &= jac > 0.9995 and got[3] == ref[3] and (ndev > 1 or same)
            # co-clustering on the reference graph
            if not dsamp:
                sets, seeds = synth.resample_sets(m.ncells, 0.75, 13, 4)
                n1, n2, cnt, ns = mg.coclust(m.ncells, ref[0], ref[1], ref[2], sets, seeds, 15, 1.05, 10)
                whole = _lib.CoClust(ctx, m.ncells, ref[0], ref[1], ref[2], sets, seeds, 15, 1.05, 10)
                w1, w2, wc, wn = whole.pairs()
                same_c = np.array_equal(n1, w1) and np.array_equal(n2, w2) and np.array_equal(cnt, wc) and np.array_equal(ns, wn)
                print(f"multi coclust devices={ndev}: pairs {n1.size} vs {w1.size}, identical {same_c}", flush=True)
                ok &= same_c
                whole.destroy()
            mg.close()
    print("MULTI_CHECK", "PASS" if ok else "FAIL", flush=True)
    return ok


def multi_process():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    ctx = _lib.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(_lib.Comm.unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    comm = _lib.Comm(ctx, world, rank, bytes(idt.numpy().tobytes()))
    m, feats, x = inputs()
    cut = [m.ncells * r // world for r in range(world + 1)]
    ok = True
    for dsamp in (False, True):
        local_csc = _lib.Csc.upload_range(ctx, m.ngenes, m.indptr, m.indices, x, cut[rank], cut[rank + 1])
        g = comm.build_sharded(local_csc, cut[rank], feats, K, KX, KB, 7.0, dsamp=dsamp, seed=5)
        res = comm.gather_edges(g, 0)
        g.destroy()
        local_csc.free()
        if rank == 0:
            ref = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, x, feats, K, KX, KB, 7.0, dsamp=dsamp, seed=5)
            jac = jaccard(res, ref)
            same = all(np.array_equal(a, b) for a, b in zip(res, ref[:3]))
            print(f"sharded build dsamp={dsamp} ranks={world}: edges {res[0].size} vs {ref[0].size}, jaccard {jac:.6f}, identical {same}", flush=True)
            ok &= jac > 0.9995
    # co-clustering: strided resamples + in-library NCCL reduce
    src = np.zeros(0, np.int32)
    if rank == 0:
        src, dst, w = ref[0], ref[1], ref[2]
    ne = torch.tensor([src.size])
    dist.broadcast(ne, 0)
    bufs = [torch.zeros(int(ne), dtype=torch.int32), torch.zeros(int(ne), dtype=torch.int32), torch.zeros(int(ne), dtype=torch.float64)]
    if rank == 0:
        bufs = [torch.from_numpy(np.ascontiguousarray(a)) for a in (src, dst, w)]
    for t in bufs:
        dist.broadcast(t, 0)
    es, ed, ew = (t.numpy() for t in bufs)
    sets, seeds = synth.resample_sets(m.ncells, 0.75, 13, 4)
    cc = _lib.CoClust(ctx, m.ncells, es, ed, ew, sets, seeds, 15, 1.05, 10, rank=rank, nranks=world)
    comm.coclust_reduce(cc, 0)
    if rank == 0:
        out = cc.pairs()
        whole = _lib.CoClust(ctx, m.ncells, es, ed, ew, sets, seeds, 15, 1.05, 10)
        exp = whole.pairs()
        same = all(np.array_equal(a, b) for a, b in zip(out, exp))
        print(f"sharded coclust ranks={world}: pairs {out[0].size} vs {exp[0].size}, identical {same}", flush=True)
        ok &= same
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    if rank == 0:
        print("MULTI_CHECK", "PASS" if ok else "FAIL", flush=True)
    cc.destroy()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    return bool(int(flag))


if __name__ == "__main__":
    sys.exit(0 if (multi_process() if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 else single_process()) else 1)
