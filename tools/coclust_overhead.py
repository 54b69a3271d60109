"""where the non-kernel time of the C3 co-clustering call goes: wall time of the constructor and of pairs() against the
library's stage timers"""
import sys, os, time
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
sets, seeds = synth.resample_sets(m.ncells, 0.75, 500, 77)
ctx.set_option("profile", 1)
for rep in range(3):
    ctx.timings_reset()
    t0 = time.perf_counter()
    cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets, seeds, 20, 1.05, 10)
    t1 = time.perf_counter()
    n1, n2, cnt, ns = cc.pairs()
    t2 = time.perf_counter()
    cc.destroy()
    t3 = time.perf_counter()
    tm = {k: round(v[0], 2) for k, v in ctx.timings().items()}
    print(f"ctor {1e3 * (t1 - t0):.1f} ms, pairs {1e3 * (t2 - t1):.1f} ms, destroy {1e3 * (t3 - t2):.1f} ms; stages {tm}; pairs {n1.size}", flush=True)
