"""one C2 step (for profiling individual kernels)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib, synth
m = synth.make_umi_matrix(100000, 1000, n_types=24, median_umis=1500.0, seed=1234)
ctx = _lib.Context(0)
csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
for it in range(2):
    g = _lib.CGraph.from_csc(ctx, csc, np.arange(1000, dtype=np.int32), 7.0)
    g.knn(100, 10); g.mutual(3); print(g.prune()); g.destroy()
