"""method = "edges" co-clustering at a size where method = "full" cannot run on one GPU's memory budget comfortably:
the C2 graph (100,000 cells, 7.5 M edges), n_resamp resamples, p = 0.75, min_mc_size = 20.  Prints the timing and checks the
per-edge counts against a recount from the per-resample assignments.  usage: python tools/edges_scale.py [n_resamp] [ncells]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth

n_resamp = int(sys.argv[1]) if len(sys.argv) > 1 else 148
ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
ctx = _lib.Context(0)
m = synth.make_umi_matrix(ncells, 1000, n_types=24, median_umis=1500.0, seed=1234)
feats = np.arange(1000, dtype=np.int32)
t0 = time.perf_counter()
src, dst, w, nv = _lib.cgraph_bknn_oneshot(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64), feats, 100, 10, 3, 7.0)
t_graph = time.perf_counter() - t0
sets, seeds = synth.resample_sets(ncells, 0.75, n_resamp, 77)
ctx.set_option("profile", 1)
out = {}
for rep in range(2):
    ctx.timings_reset()
    t0 = time.perf_counter()
    cc = _lib.CoClust(ctx, ncells, src, dst, w, sets, seeds, 20, 1.05, 10, knn=int(round(src.size / ncells)), method="edges")
    e1, e2, ec, ns = cc.edge_counts()
    dt = time.perf_counter() - t0
    st = {k: round(v[0], 2) for k, v in ctx.timings().items()}
    if rep == 1:
        mc, sw = cc.assignments()
        chk = min(n_resamp, 8)
        # recount on a subset of the resamples is a lower bound; the full recount must match exactly
        want = np.zeros(e1.size, np.int64)
        for r in range(n_resamp):
            a = mc[r][e1]
            want += (a >= 0) & (a == mc[r][e2])
        ok = bool(np.array_equal(want, ec)) and int(ns.sum()) == n_resamp * sets.shape[1]
        out = {"config": f"method=edges: {ncells} cells, {src.size} edges, {n_resamp} resamples, p=0.75, min_mc_size=20, knn={int(round(src.size / ncells))}",
               "graph_s": round(t_graph, 3), "coclust_s": round(dt, 3), "resamples_per_s": round(n_resamp / dt, 1), "stages_ms": st,
               "mean_sweeps": float(np.mean(sw)), "edges_with_count": int((ec > 0).sum()), "max_count": int(ec.max()),
               "clusters_first_resample": int(mc[0].max() + 1), "recount_from_assignments_equal": ok}
    cc.destroy()
print(json.dumps(out), flush=True)
sys.exit(0 if out.get("recount_from_assignments_equal") else 1)
