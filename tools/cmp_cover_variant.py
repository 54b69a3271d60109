"""bit-exactness of an experimental cover-kernel variant (MCB_COVER_KERNEL) against the default on the C1 graph.
usage: python tools/cmp_cover_variant.py <variant> [<variant> ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
ctx = _lib.Context(0)
m, feats, prm = synth.config("C1")
src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, prm["K"], 10, 3, 7.0, dsamp=prm["dsamp"], seed=42)
sets, seeds = synth.resample_sets(m.ncells, 0.75, 40, 77)
out = {}
for k in ["auto"] + sys.argv[1:]:
    os.environ["MCB_COVER_KERNEL"] = k
    cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets, seeds, 20, 1.05, 10)
    out[k] = cc.assignments()
    cc.destroy()
    if k != "auto":
        print(k, "identical to auto:", all(np.array_equal(a, b) for a, b in zip(out["auto"], out[k])))
