"""diagnostic: which bench ingredient slows the step"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
variant = sys.argv[1]
from metacell_b200 import _lib, synth
m = synth.make_umi_matrix(100000, 1000, n_types=24, median_umis=1500.0, seed=1234)
keep = []
if variant != "plain":
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
ctx = _lib.Context(0)
if variant in ("wrap", "pinned"):
    dp = torch.from_numpy(m.indptr.astype(np.int64)).cuda(); di = torch.from_numpy(m.indices.astype(np.int32)).cuda()
    dx = torch.from_numpy(m.data.astype(np.int64)).cuda().to(torch.int32)
    csc = _lib.Csc.wrap_device(ctx, m.ngenes, m.ncells, m.nnz, dp.data_ptr(), di.data_ptr(), dx.data_ptr(), keep=(dp, di, dx))
else:
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
if variant == "pinned":
    keep.append(torch.empty(10000000, dtype=torch.float64).pin_memory())
    keep.append(torch.from_numpy(m.data.astype(np.float64)).pin_memory())
feats = np.arange(1000, dtype=np.int32)
for it in range(4):
    ts = [time.perf_counter()]
    g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0); ts.append(time.perf_counter())
    g.knn(100, 10); ts.append(time.perf_counter())
    g.mutual(3); ts.append(time.perf_counter())
    g.prune(); ts.append(time.perf_counter())
    g.destroy(); ctx.sync(); ts.append(time.perf_counter())
    print(variant, it, ["%.1f" % ((b - a) * 1e3) for a, b in zip(ts, ts[1:])], "total %.1f ms" % ((ts[-1] - ts[0]) * 1e3), flush=True)
