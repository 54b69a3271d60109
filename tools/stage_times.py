"""Per-stage CUDA-event timings of the C2-shaped balanced-graph build (device-resident input), for quick A/B runs of a
kernel change or a tuning environment variable.  usage: [ENV=..] python tools/stage_times.py [--cells N] [--genes G] [--K K] [--opt key=value ...]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=1000)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--passes", type=int, default=4)
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    from metacell_b200 import _lib, synth
    m = synth.make_umi_matrix(a.cells, a.genes, n_types=24, median_umis=1500.0, seed=1234)
    feats = np.arange(a.genes, dtype=np.int32)
    ctx = _lib.Context(0)
    for kv in a.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    acc = {}
    for it in range(a.passes):
        ctx.set_option("profile", 1 if it else 0)
        ctx.timings_reset()
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        g.knn(a.K, 10)
        g.mutual(3)
        ne = g.prune()
        if it:
            for k, v in ctx.timings().items():
                acc.setdefault(k, []).append(v[0])
        stats = g.knn_stats()
        g.destroy()
    st = {k: round(float(np.median(v)), 3) for k, v in acc.items()}
    print(json.dumps({"opts": a.opt, "env": {k: v for k, v in os.environ.items() if k.startswith("MCB_")}, "edges": ne, "knn": stats,
                      "sum_ms": round(sum(st.values()), 3), "stages_ms": st}))


if __name__ == "__main__":
    main()
