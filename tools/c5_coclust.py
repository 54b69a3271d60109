"""C5-shaped co-clustering (~20,000 cells, low UMI/cell, K = 150, 1,000 resamples): graph on the GPU, then
the resampled cover with each kernel family.  usage: python tools/c5_coclust.py [n_resamp]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
n_resamp = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = _lib.Context(0)
m = synth.make_umi_matrix(20000, 800, n_types=40, median_umis=800.0, seed=55)
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
g = _lib.CGraph.from_csc(ctx, csc, np.arange(800, dtype=np.int32), 7.0)
g.knn(150, 10); g.mutual(3); g.prune()
src, dst, w = g.edges()
N = m.ncells
print(f"C5 graph: {g.M} valid cells, {src.size} edges, max out {np.bincount(src).max()}, max in {np.bincount(dst).max()}", flush=True)
g.destroy(); csc.free()
sets, seeds = synth.resample_sets(N, 0.75, n_resamp, 77)
ctx.set_option("profile", 1)
for kern in (sys.argv[2:] or ["auto", "warp16", "fast"]):
    os.environ["MCB_COVER_KERNEL"] = kern
    for rep in range(2):
        ctx.timings_reset()
        t0 = time.time()
        cc = _lib.CoClust(ctx, N, src, dst, w, sets, seeds, 20, 1.05, 10)
        n1 = cc.pairs()[0].size
        wall = time.time() - t0
        tm = {k: round(v[0], 2) for k, v in ctx.timings().items()}
        _, sw = cc.assignments()
        cc.destroy()
    print(f"{kern:8s}: wall {wall * 1e3:8.1f} ms -> {n_resamp / wall:7.0f} resamples/s  stages {tm}  pairs {n1}  mean sweeps {sw.mean():.1f}", flush=True)
