"""Generates tests/golden/oracle_small.npz: outputs of the CPU oracle on a tiny seeded input.

PARITY UNPINNED: the reference (tanaylab/metacell) ships no tests, fixtures or golden vectors for this
path and its arithmetic lives in the absent tgstat package, so these vectors pin
the *oracle* against regressions, not the oracle against tgstat.  The only external known answers are
the Random123 Philox4x32-10 KAT vectors checked in tests/test_oracle.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from metacell_b200 import synth          # noqa: E402
from oracle import oracle as O           # noqa: E402


def main():
    m = synth.make_umi_matrix(240, 120, n_types=4, median_umis=900.0, seed=99)
    feats = np.arange(0, 120, 2, dtype=np.int32)
    tot = m.col_sums()
    depth = O.downsamp_depth(tot)
    ds, kept = O.downsample(m.indptr, m.data, 600.0, 42)
    fm = np.full(m.ngenes, -1, np.int32)
    fm[feats] = np.arange(feats.size)
    Z, valid = O.features(m.indptr, m.indices, m.data, fm, feats.size)
    col, cor = O.cor_knn(Z[valid], 40)
    deg, dst, s = O.balance(col, 8, 5, 3)
    src_e, dst_e, w_e = O.edges_from_balance(deg, dst, 8)
    M = int(valid.sum())
    sets, seeds = synth.resample_sets(M, 0.75, 3, 5)
    csr = O.graph_to_csr(M, src_e, dst_e, w_e)
    mcs = np.stack([O.graph_cover(M, csr, sets[r], 5, 1.05, 10, int(seeds[r]))[0] for r in range(3)])
    cnt, nsamp, sweeps = O.coclust(M, csr, sets, seeds, 5, 1.05, 10)
    n1, n2 = np.nonzero(cnt)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz")
    np.savez_compressed(out, indptr=m.indptr, indices=m.indices, data=m.data, ngenes=m.ngenes, feats=feats,
                        depth=depth, ds=ds, kept=kept, Z=Z.astype(np.float64), valid=valid, knn_col=col, knn_cor=cor,
                        deg=deg, bal_dst=dst, bal_s=s, sets=sets, seeds=seeds, mcs=mcs, sweeps=sweeps,
                        pair1=n1.astype(np.int32), pair2=n2.astype(np.int32), pair_cnt=cnt[n1, n2], nsamp=nsamp)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
