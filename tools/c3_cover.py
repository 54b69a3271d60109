"""C3 co-clustering (8,000 cells, K = 100 graph, 500 resamples) once or twice, for ncu captures of the cover kernel.
usage: python tools/c3_cover.py [n_resamp] [reps]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
n_resamp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _lib.Context(0)
m, feats, prm = synth.config("C1")
src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, prm["K"], 10, 3, 7.0, dsamp=prm["dsamp"], seed=42)
sets, seeds = synth.resample_sets(m.ncells, 0.75, n_resamp, 77)
ctx.set_option("profile", 1)
for rep in range(reps):
    ctx.timings_reset()
    t0 = time.perf_counter()
    cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets, seeds, 20, 1.05, 10)
    n1, n2, cnt, ns = cc.pairs()
    t1 = time.perf_counter()
    _, sw = cc.assignments()
    cc.destroy()
    print(f"{n_resamp} resamples: {1e3 * (t1 - t0):.1f} ms -> {n_resamp / (t1 - t0):.0f} /s; stages", {k: round(v[0], 2) for k, v in ctx.timings().items()}, "mean sweeps", float(np.mean(sw)), flush=True)
