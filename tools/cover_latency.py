"""diagnostic: graph-cover time vs number of concurrent resamples (latency- or throughput-bound?)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
ctx = _lib.Context(0)
m = synth.make_umi_matrix(8000, 600, n_types=12, median_umis=3000.0, seed=1239)
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
g = _lib.CGraph.from_csc(ctx, csc, np.arange(600, dtype=np.int32), 7.0)
g.knn(100, 10); g.mutual(3); g.prune()
src, dst, w = g.edges()
g.destroy(); csc.free()
sets, seeds = synth.resample_sets(m.ncells, 0.75, 1184, 77)
ctx.set_option("profile", 1)
for n in (1, 37, 148, 296, 500, 592, 1184):
    for rep in range(2):
        ctx.timings_reset()
        cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets[:n], seeds[:n], 20, 1.05, 10)
        t = ctx.timings()["graph_cover"][0]
        _, sw = cc.assignments()
        cc.destroy()
    print(f"resamples {n:5d}: graph_cover {t:8.2f} ms  mean sweeps {sw.mean():.1f} max {sw.max()}  -> {n / t * 1e3:.0f} resamples/s", flush=True)
