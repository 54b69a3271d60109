"""Multi-GPU plumbing for the two sharded stages: one process per GPU,
`torch.distributed` (NCCL over NVLink) only for the exchange steps the path really has.

  balanced cgraph : rows of the kNN stage are sharded contiguously across ranks; the path has two
                    real exchanges -- all-gather of the kNN column lists (every rank needs rank(j,i)
                    of foreign rows) and all-gather of the per-target in-degree thresholds.
  co-clustering   : resamples are strided across ranks with the graph replicated; one reduce(sum)
                    of the co-occurrence counters and of samples[N] to rank 0.

PyTorch appears here for process groups and for viewing library-owned device buffers as tensors;
it never computes anything on the path.  The stage engine is an object with the methods of
`CudaCGraphEngine`, so the same plumbing is exercised on CPU (gloo, world_size 2) by the tests with
an engine built on the oracle.
"""
import numpy as np
import torch
import torch.distributed as dist


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_view(ptr, nbytes, dtype, device):
    """torch view (no copy) of a library-owned device buffer"""
    t = torch.as_tensor(_DevMem(ptr, nbytes), device=torch.device("cuda", device))
    return t.view(dtype)


def _all_gather_blocks(full, rank, world, group=None):
    """in-place all-gather: `full` is [world * blk]; every rank has filled its own block"""
    blk = full.numel() // world
    mine = full[rank * blk:(rank + 1) * blk].clone()
    if full.is_cuda:
        dist.all_gather_into_tensor(full, mine, group=group)
    else:
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine, group=group)
        for r in range(world):
            full[r * blk:(r + 1) * blk].copy_(outs[r])


class CudaCGraphEngine:
    """Stage engine over a metacell_b200._lib.CGraph (the product path)."""

    def __init__(self, cgraph):
        self.g = cgraph
        self.device = cgraph.ctx.device

    def knn(self, K, k_expand, rank, world):
        self.g.knn(K, k_expand, rank, world)

    # staged kNN (symmetric schedule): thresholds, record exchange, finish
    def knn_begin(self, K, k_expand, rank, world):
        self.g.knn_begin(K, k_expand, rank, world)

    def knn_thr_full(self, world):
        ptr, rpr = self.g.knn_thr_buffer()
        return device_view(ptr, world * rpr * 4, torch.float32, self.device)

    def knn_filter(self):
        return self.g.knn_filter()

    def records_view(self, ptr, n_records):
        return device_view(ptr, max(n_records, 1) * 16, torch.int32, self.device)[: n_records * 4]

    def knn_recv_buffer(self, n):
        return self.g.knn_recv_buffer(n)

    def knn_finish(self, n):
        self.g.knn_finish(n)

    def knn_full(self, world):
        ptr, rpr, kc = self.g.knn_buffer()
        return device_view(ptr, world * rpr * kc * 4, torch.int32, self.device)

    def mutual(self, k_beta):
        self.g.mutual(k_beta)

    def thr_full(self, world):
        ptr, rpr = self.g.thr_buffer()
        return device_view(ptr, world * rpr * 8, torch.int64, self.device)

    def prune(self):
        return self.g.prune()

    def edges(self):
        return self.g.edges()

    def sync(self):
        self.g.ctx.sync()
        torch.cuda.synchronize(self.device)


def _knn_symmetric(engine, K, k_expand, rank, world, group=None):
    """kNN stage with every unordered cell pair contracted once across all ranks: each rank takes a share of
    the symmetric tile list and produces candidate records for rows of every rank; the records travel to
    their owners in one NCCL all-to-all.  Exchanges: thresholds (all-gather), record counts (all-gather),
    records (all-to-all)."""
    engine.knn_begin(K, k_expand, rank, world)
    thr = engine.knn_thr_full(world)
    engine.sync()
    _all_gather_blocks(thr, rank, world, group)
    engine.sync()
    send_ptr, counts = engine.knn_filter()
    dev = thr.device
    mine = torch.tensor(counts, dtype=torch.int64, device=dev)
    allc = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine, group=group)
    recv_counts = allc.view(world, world)[:, rank].tolist()
    n_send, n_recv = int(sum(counts)), int(sum(recv_counts))
    recv_ptr = engine.knn_recv_buffer(n_recv)
    send_t = engine.records_view(send_ptr, n_send)
    recv_t = engine.records_view(recv_ptr, n_recv)
    engine.sync()
    dist.all_to_all_single(recv_t, send_t, output_split_sizes=[int(c) * 4 for c in recv_counts],
                           input_split_sizes=[int(c) * 4 for c in counts], group=group)
    engine.sync()
    engine.knn_finish(n_recv)


def cgraph_sharded(engine, K, k_expand, k_beta, rank, world, group=None, fetch_edges=True, symmetric=True):
    """Runs A5..A7 with rows sharded over `world` ranks.  Returns this rank's (src, dst, w), or the edge
    count when fetch_edges is False (edges stay on the device).  symmetric=False keeps the exchange-free
    rectangular schedule (each rank contracts its rows against all columns: twice the tensor work)."""
    if world > 1 and symmetric and hasattr(engine, "knn_begin"):
        _knn_symmetric(engine, K, k_expand, rank, world, group)
    else:
        engine.knn(K, k_expand, rank, world)
    if world > 1:
        full = engine.knn_full(world)
        engine.sync()
        _all_gather_blocks(full, rank, world, group)
        engine.sync()
    engine.mutual(k_beta)
    if world > 1:
        full = engine.thr_full(world)
        engine.sync()
        _all_gather_blocks(full, rank, world, group)
        engine.sync()
    n = engine.prune()
    return engine.edges() if fetch_edges else n


def gather_edges(src, dst, w, rank, world, device=None, group=None):
    """Concatenate every rank's edge rows on rank 0 (rank order == source order).  Others get None."""
    if world == 1:
        return src, dst, w
    dev = torch.device("cuda", device) if device is not None else torch.device("cpu")
    n = torch.tensor([src.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)

    def _gather(a, dtype):
        t = torch.zeros(mx, dtype=dtype, device=dev)
        t[:a.size] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        outs = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, outs, dst=0, group=group)
        if rank != 0:
            return None
        return np.concatenate([outs[r][:sizes[r]].cpu().numpy() for r in range(world)])

    s, d, ww = _gather(src, torch.int32), _gather(dst, torch.int32), _gather(w, torch.float64)
    return (s, d, ww) if rank == 0 else None


def coclust_sharded(coclust, rank, world, group=None):
    """Reduce(sum) this rank's co-occurrence counters and sample counts into rank 0, then extract the
    (node1, node2, cnt) rows there.  `coclust` is a metacell_b200._lib.CoClust built with (rank, world)."""
    if world > 1:
        pc, pn = coclust.buffers()
        dev = coclust.ctx.device
        N = coclust.N
        cnt = device_view(pc, N * N * 4, torch.int32, dev)
        ns = device_view(pn, N * 4, torch.int32, dev)
        coclust.ctx.sync()
        torch.cuda.synchronize(dev)
        dist.reduce(cnt, dst=0, op=dist.ReduceOp.SUM, group=group)
        dist.reduce(ns, dst=0, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.synchronize(dev)
    if rank == 0:
        return coclust.pairs()
    return None
