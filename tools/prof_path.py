"""One pass of the whole path (upload -> col sums -> downsample -> features -> kNN -> balancing -> edges) on one B200,
for `ncu` captures of the HBM-bound kernels (downsample, features, record binning, ranked lists, balancing).

usage: python tools/prof_path.py [--cells 100000] [--genes 2000] [--feats 1000] [--K 100] [--passes 2]
Prints the per-stage CUDA-event timings of the last pass.  The matrix has more gene rows than feature genes so
that the down-sampling kernel sees a realistic nnz (all genes x cells).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--feats", type=int, default=1000)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--median-umis", type=float, default=3000.0)
    a = ap.parse_args()
    from metacell_b200 import _lib, synth
    m = synth.make_umi_matrix(a.cells, a.genes, n_types=24, median_umis=a.median_umis, seed=1234)
    feats = np.sort(np.random.default_rng(1).choice(a.genes, a.feats, replace=False)).astype(np.int32)
    ctx = _lib.Context(0)
    ctx.set_option("profile", 1)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    out = {}
    for it in range(a.passes):
        ctx.timings_reset()
        tot = csc.col_sums()
        depth = _lib.downsamp_depth(tot)
        ds = csc.downsample(depth, 42, add_non_dsamp=True)
        g = _lib.CGraph.from_csc(ctx, ds, feats, 7.0)
        g.knn(a.K, 10)
        g.mutual(3)
        ne = g.prune()
        st = ctx.timings()
        out = {"cells": a.cells, "genes": a.genes, "feats": a.feats, "nnz": m.nnz, "umis": int(tot.sum()), "depth": depth,
               "edges": ne, "knn": g.knn_stats(), "stages_ms": {k: round(v[0], 3) for k, v in st.items()}}
        g.destroy()
        ds.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
