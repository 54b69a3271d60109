"""wall-clock per API call for a few consecutive steps (diagnostic)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = synth.make_umi_matrix(n, 1000, n_types=24, median_umis=1500.0, seed=1234)
ctx = _lib.Context(0)
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
feats = np.arange(1000, dtype=np.int32)
for it in range(5):
    ts = [time.perf_counter()]
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0); ctx.sync(); ts.append(time.perf_counter())
    g.knn(100, 10); ctx.sync(); ts.append(time.perf_counter())
    g.mutual(3); ctx.sync(); ts.append(time.perf_counter())
    g.prune(); ctx.sync(); ts.append(time.perf_counter())
    g.destroy(); ctx.sync(); ts.append(time.perf_counter())
    print(it, ["%.1f" % ((b - a) * 1e3) for a, b in zip(ts, ts[1:])], "total %.1f ms" % ((ts[-1] - ts[0]) * 1e3), flush=True)
