"""World-size-2 (and 3) gloo tests of the multi-rank PROTOCOL on CPU (metacell_b200/dist.py: the numpy mirror of what
libmcb200.so runs over NCCL).  No GPU here, so every stage's compute comes from the CPU oracle (test infrastructure);
what is under test is the rank-level logic the GPU path shares: contiguous cell / row splits, variable-block
all-gathers, the routed balancing join with its symmetry trick, the all-gather of the in-degree cuts, the gather of the
edge rows, the strided resamples + reduce -- and that the sharded result equals the single-rank oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, KX, KB = 6, 5, 3


def _matrix():
    from metacell_b200 import synth
    m = synth.make_umi_matrix(333, 150, n_types=5, median_umis=900.0, seed=21)       # odd size: ragged last block
    data = m.data.copy()
    for c in (0, 170, 332):                                                            # degenerate cells on both ranks
        data[m.indptr[c]:m.indptr[c + 1]] = 0
    return m, data


class OracleEngine:
    def __init__(self):
        from oracle import oracle as O
        self.O = O
        self.m, self.data = _matrix()

    def features(self, cells):
        c0, c1 = cells
        m = self.m
        p = m.indptr[c0:c1 + 1] - m.indptr[c0]
        sl = slice(m.indptr[c0], m.indptr[c1])
        Z, valid = self.O.features(p, m.indices[sl], self.data[sl], np.arange(150, dtype=np.int32), 150)
        return Z[valid], np.arange(c0, c1)[valid]

    def knn(self, Z, row0, row1, Kc):
        col, _ = self.O.cor_knn(Z, Kc, row0, row1)
        return col


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metacell_b200 import dist as mdist
    from metacell_b200 import synth
    from oracle import oracle as O
    eng = OracleEngine()
    n = eng.m.ncells
    cells = (n * rank // world, n * (rank + 1) // world)
    src, dst, w = mdist.cgraph_sharded(eng, cells, K, KX, KB)
    res = mdist.gather_edges(src, dst, w)
    # co-clustering: resamples strided over the ranks on the gathered graph (broadcast from rank 0), counters reduced
    obj = [res]
    dist.broadcast_object_list(obj, src=0)
    es, ed, ew = obj[0]
    sets, seeds = synth.resample_sets(n, 0.75, 7, 5)
    mine = list(range(rank, 7, world))
    csr = O.graph_to_csr(n, es, ed, ew)
    cnt, ns, _ = O.coclust(n, csr, sets[mine], seeds[mine], 8, 1.05, 10) if mine else (np.zeros((n, n), np.uint32), np.zeros(n, np.int32), None)
    red = mdist.coclust_sharded(cnt, ns)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), src=res[0], dst=res[1], w=res[2], cnt=red[0], ns=red[1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_equals_single_rank(tmp_path, world):
    from metacell_b200 import synth
    from oracle import oracle as O
    port = 29500 + (os.getpid() % 1000) + world
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    got = np.load(tmp_path / "out.npz")
    m, data = _matrix()
    Z, valid = O.features(m.indptr, m.indices, data, np.arange(150, dtype=np.int32), 150)
    assert not valid[[0, 170, 332]].any()
    cells = np.nonzero(valid)[0]
    col, _ = O.cor_knn(Z[valid], K * KX)
    deg, dst, _ = O.balance(col, K, KX, KB)
    src, d, w = O.edges_from_balance(deg, dst, K)
    assert np.array_equal(got["src"], cells[src]) and np.array_equal(got["dst"], cells[d]) and np.array_equal(got["w"], w)
    sets, seeds = synth.resample_sets(m.ncells, 0.75, 7, 5)
    csr = O.graph_to_csr(m.ncells, got["src"], got["dst"], got["w"])
    cnt, ns, _ = O.coclust(m.ncells, csr, sets, seeds, 8, 1.05, 10)
    assert np.array_equal(got["cnt"], cnt) and np.array_equal(got["ns"], ns)


def test_resample_striding_partitions_the_work():
    """rank r takes resamples r, r + world, ... (mcb_coclust_run contract): a partition for any world size"""
    for world in (1, 2, 3, 8):
        seen = sorted(r for rank in range(world) for r in range(rank, 500, world))
        assert seen == list(range(500))


def test_row_split_covers_every_row_once():
    from metacell_b200.dist import row_split
    for M, world in ((10, 3), (333, 2), (7, 8), (100000, 8)):
        rpr, split = row_split(M, world)
        rows = [r for a, b in split for r in range(a, b)] if M < 1000 else None
        assert split[0][0] == 0 and split[-1][1] == M and all(split[q][1] == split[q + 1][0] for q in range(world - 1))
        if rows is not None:
            assert rows == list(range(M))
