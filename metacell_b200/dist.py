"""Protocol mirror of the multi-rank build -- TEST HARNESS ONLY (tests/test_dist_gloo.py).

The product path runs its exchanges inside libmcb200.so over NCCL (csrc/comm.cu, api.cu: mcb_cgraph_build_sharded,
mcb_coclust_reduce; csrc/multi.cu for one process driving several devices) and imports neither torch nor this
module.  What lives here is the same RANK-LEVEL PROTOCOL written with numpy over `torch.distributed` (gloo on CPU,
world_size 2), so that the sharding rules and the exchanges -- above all the routed, symmetric balancing join -- can be
checked against the single-rank oracle on a box without GPUs:

  ingest     contiguous cell ranges per rank; valid-row counts all-gathered; feature rows all-gathered (variable blocks)
  kNN        rank r ranks the rows [r * rpr, (r+1) * rpr), rpr = ceil(M / world); the lists stay with their owner
  join       every list entry (i -> j, rank r) is routed to the owner of row j (all-to-all of 8-byte records); the owner
             finds rank(j, i) in ITS OWN list of j and, because s(i,j) = rank(i,j) * rank(j,i) = s(j,i), records the
             score in row j's own array at position rank(j,i): nothing travels back
  cuts       per-target in-degree cut from each row's own scores; all-gather (8 bytes per row)
  prune      out-degree cut of the owned rows; edges gathered to rank 0 in rank (= source) order
  coclust    resamples strided over the ranks, graph replicated, reduce(sum) of the counters to rank 0

The compute of each stage is supplied by an engine object (the tests use one built on the CPU oracle).
"""
import numpy as np
import torch
import torch.distributed as dist


def _allgather_var(x, group=None):
    """all-gather of numpy blocks whose first dimension differs per rank -> list of blocks in rank order"""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(max(sizes), 1)
    pad = np.zeros((mx,) + x.shape[1:], dtype=x.dtype)
    pad[:x.shape[0]] = x
    outs = [torch.zeros(pad.shape, dtype=torch.from_numpy(pad).dtype) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(pad), group=group)
    return [outs[r][:sizes[r]].numpy() for r in range(world)]


def _alltoall_var(blocks, group=None):
    """blocks[r] (numpy, 1-D) goes to rank r; returns the list of blocks received, by sender"""
    world = dist.get_world_size(group)
    got = []
    for dst_rank in range(world):           # gloo has no all_to_all: one gather per destination
        parts = _gather_var(blocks[dst_rank], dst_rank, group)
        if parts is not None:
            got = parts
    return got


def _gather_var(x, root, group=None):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    mx = max(max(sizes), 1)
    pad = np.zeros(mx, dtype=x.dtype)
    pad[:x.shape[0]] = x
    t = torch.from_numpy(pad)
    outs = [torch.zeros_like(t) for _ in range(world)] if rank == root else None
    dist.gather(t, outs, dst=root, group=group)
    if rank != root:
        return None
    return [outs[r][:sizes[r]].numpy().copy() for r in range(world)]


def row_split(M, world):
    """(rows_per_rank, [(row0, row1)] per rank): the contiguous split the library uses (ceil_div)"""
    rpr = -(-M // world)
    return rpr, [(min(M, r * rpr), min(M, (r + 1) * rpr)) for r in range(world)]


def routed_join(col_own, row0, M, world, K, k_expand, k_beta, group=None, thresh_strict=True, keep_self=False):
    """The balancing join + in-degree cuts for the owned rows, as csrc/balance.cu + api.cu::cgraph_mutual run it.
    col_own: [rows][Kc] kNN lists of the rows [row0, row0 + rows).  Returns (sym [rows][Kc] uint32, thr_in [M] uint64)."""
    rank = dist.get_rank(group)
    rows, Kc = col_own.shape
    rpr, _ = row_split(M, world)
    smax = K * K * k_expand
    # ---- records of the own entries, routed by the owner of the target
    i = np.repeat(np.arange(row0, row0 + rows, dtype=np.int64), Kc)
    r = np.tile(np.arange(Kc, dtype=np.int64), rows)
    j = col_own.reshape(-1).astype(np.int64)
    ok = j >= 0
    if not keep_self:
        ok &= j != i
    i, j, r = i[ok], j[ok], r[ok]
    owner = j // rpr
    send = [np.stack([i[owner == q], j[owner == q], r[owner == q]], axis=1).reshape(-1) for q in range(world)]
    recv = _alltoall_var(send, group)
    rec = np.concatenate([b.reshape(-1, 3) for b in recv]) if recv else np.zeros((0, 3), np.int64)
    # ---- probe: rank(j, i) from this rank's own list of j; the score goes into row j's array (symmetry)
    sym = np.zeros((rows, Kc), dtype=np.uint32)
    if rec.shape[0] and rows:
        key_own = (np.repeat(np.arange(rows, dtype=np.int64), Kc) * M + col_own.reshape(-1).astype(np.int64))
        valid = col_own.reshape(-1) >= 0
        order = np.argsort(key_own[valid], kind="stable")
        keys_sorted = key_own[valid][order]
        pos_sorted = np.nonzero(valid)[0][order]                    # flat position (lrow * Kc + rank0) of each key
        q = (rec[:, 1] - row0) * M + rec[:, 0]                      # (row of j, column i)
        at = np.searchsorted(keys_sorted, q)
        hit = (at < keys_sorted.size) & (keys_sorted[np.minimum(at, keys_sorted.size - 1)] == q)
        flat = pos_sorted[at[hit]]
        rji = flat % Kc + 1
        s = (rec[hit, 2] + 1) * rji
        keep = (s < smax) if thresh_strict else (s <= smax)
        sym.reshape(-1)[flat[keep]] = s[keep].astype(np.uint32)
    # ---- in-degree cut of every owned row from its own scores (key = (s, other node)), then all-gather
    kin = K * k_beta
    thr_own = np.full(rows, np.iinfo(np.uint64).max, dtype=np.uint64)
    for lrow in range(rows):
        m = sym[lrow] > 0
        if m.sum() > kin:
            keys = (sym[lrow][m].astype(np.uint64) << np.uint64(32)) | col_own[lrow][m].astype(np.uint64)
            thr_own[lrow] = np.sort(keys)[kin - 1]
    thr = np.concatenate(_allgather_var(thr_own.view(np.int64), group)).view(np.uint64)
    return sym, thr


def prune_rows(col_own, sym, thr, row0, K):
    """out-degree cut of the owned rows (csrc/balance.cu prune_kernel): survivors of the target's in-degree cut, best K
    by (s, target).  Returns (src, dst, w) of the owned rows, ordered by (source, place)."""
    src, dst, w = [], [], []
    for lrow in range(col_own.shape[0]):
        i = row0 + lrow
        m = sym[lrow] > 0
        j = col_own[lrow][m].astype(np.uint64)
        s = sym[lrow][m].astype(np.uint64)
        inkey = (s << np.uint64(32)) | np.uint64(i)
        ok = inkey <= thr[j.astype(np.int64)]
        keys = np.sort((s[ok] << np.uint64(32)) | j[ok])[:K]
        src += [i] * keys.size
        dst += (keys & np.uint64(0xffffffff)).astype(np.int64).tolist()
        w += (1.0 - np.arange(keys.size) / float(K)).tolist()
    return np.asarray(src, np.int32), np.asarray(dst, np.int32), np.asarray(w, np.float64)


def cgraph_sharded(engine, cells, K, k_expand, k_beta, group=None):
    """A3..A7 over the ranks of `group`.  engine.features(cells) -> (Z_local_valid [m][G], cell ids [m]) for this rank's
    cell range; engine.knn(Z_all, row0, row1, Kc) -> kNN lists of those rows.  Returns this rank's (src, dst, w) in
    cell ids."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Zl, ids = engine.features(cells)
    Z = np.concatenate(_allgather_var(Zl, group))                   # ingest: all-gather of the feature rows
    cell_of_row = np.concatenate(_allgather_var(ids.astype(np.int64), group))
    M = Z.shape[0]
    _, split = row_split(M, world)
    row0, row1 = split[rank]
    col_own = engine.knn(Z, row0, row1, K * k_expand)
    sym, thr = routed_join(col_own, row0, M, world, K, k_expand, k_beta, group)
    src, dst, w = prune_rows(col_own, sym, thr, row0, K)
    return cell_of_row[src].astype(np.int32), cell_of_row[dst].astype(np.int32), w


def gather_edges(src, dst, w, group=None):
    """every rank's edge rows on rank 0, rank order == source order; None elsewhere"""
    parts = [_gather_var(np.ascontiguousarray(a), 0, group) for a in (src, dst, w)]
    if dist.get_rank(group) != 0:
        return None
    return tuple(np.concatenate(p) for p in parts)


def coclust_sharded(cnt_local, nsamp_local, group=None):
    """reduce(sum) of this rank's co-occurrence counters and sample counts into rank 0"""
    c = torch.from_numpy(np.ascontiguousarray(cnt_local).astype(np.int64))
    n = torch.from_numpy(np.ascontiguousarray(nsamp_local).astype(np.int64))
    dist.reduce(c, dst=0, op=dist.ReduceOp.SUM, group=group)
    dist.reduce(n, dst=0, op=dist.ReduceOp.SUM, group=group)
    if dist.get_rank(group) != 0:
        return None
    return c.numpy(), n.numpy()
