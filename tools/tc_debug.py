"""diagnostic: filter-GEMM time with parts of the epilogue disabled (MCB_TC_DEBUG) for both kernels"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    from metacell_b200 import _lib, synth
    impl = int(sys.argv[1])
    m = synth.make_umi_matrix(100000, 1000, n_types=24, median_umis=1500.0, seed=1234)
    ctx = _lib.Context(0)
    ctx.set_option("gemm_impl", impl)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    feats = np.arange(1000, dtype=np.int32)
    for it in range(3):
        if it == 2:
            ctx.set_option("profile", 1); ctx.timings_reset()
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        try:
            g.knn(100, 10)
        except Exception as e:
            pass
        g.destroy()
    t = ctx.timings()
    print("impl", impl, "dbg", os.environ.get("MCB_TC_DEBUG", "0"), "filter gemm ms", round(t["cor_filter_gemm"][0], 2), flush=True)
else:
    for impl in (0, 2):
        for dbg in ("0", "1", "2"):
            env = dict(os.environ, MCB_TC_DEBUG=dbg)
            subprocess.run([sys.executable, __file__, str(impl)], env=env)
