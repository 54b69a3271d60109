"""graph-cover time of the cover-kernel families on the C1 graph (8,000 cells, K = 100) at 1 / 62 / 500 resamples, with the
assignments of every variant compared against the first.  usage: python tools/cover_variants_time.py [variant ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
ctx = _lib.Context(0)
m, feats, prm = synth.config("C1")
src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, prm["K"], 10, 3, 7.0, dsamp=prm["dsamp"], seed=42)
sets, seeds = synth.resample_sets(m.ncells, 0.75, 500, 77)
ctx.set_option("profile", 1)
variants = sys.argv[1:] or ["batch", "incr8", "incr16", "auto"]
for n in (1, 62, 500):
    ref = None
    for k in variants:
        os.environ["MCB_COVER_KERNEL"] = ("batch2" if n > 296 else "batch4") if k == "batch" else k
        for rep in range(2):
            ctx.timings_reset()
            cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets[:n], seeds[:n], 20, 1.05, 10)
            t = ctx.timings()["graph_cover"][0]
            res = cc.assignments()
            cc.destroy()
        same = True if ref is None else all(np.array_equal(a, b) for a, b in zip(ref, res))
        ref = ref or res
        print(f"resamples {n:4d} {k:7s}: graph_cover {t:8.2f} ms -> {n / t * 1e3:8.0f} resamples/s  mean sweeps {res[1].mean():.1f}  identical to first: {same}", flush=True)
