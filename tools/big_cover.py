"""tgs_graph_cover on a 100,000-node graph (the size of the C2 cell graph; ~6,000 clusters at min size 8): time of one cover
with the kernel launch_graph_cover picks (the general kernel at this size).  usage: python tools/big_cover.py [variant ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metacell_b200 import _lib
N, min_size = 100000, 8
rng = np.random.default_rng(8)
base = np.arange(N, dtype=np.int64)
blk = base // 64 * 64
src = np.concatenate([base] * 12)
dst = np.concatenate([(base + 1) % N, (base + 2) % N] + [blk + rng.integers(0, 64, N) for _ in range(10)]) % N
keep = src != dst
_, first = np.unique(src[keep] * N + dst[keep], return_index=True)
src, dst = src[keep][first].astype(np.int32), dst[keep][first].astype(np.int32)
w = rng.random(src.size) * 0.9 + 0.1
ctx = _lib.Context(0)
ctx.set_option("profile", 1)
ref = None
for k in sys.argv[1:] or ["auto"]:
    os.environ["MCB_COVER_KERNEL"] = k
    for rep in range(2):
        ctx.timings_reset()
        t0 = time.perf_counter()
        cl, sw = _lib.graph_cover(ctx, N, src, dst, w, min_size, 1.05, 10, 99)
        wall = time.perf_counter() - t0
        t = ctx.timings()["graph_cover"][0]
    same = True if ref is None else bool(np.array_equal(ref, cl))
    ref = cl if ref is None else ref
    print(f"{k:8s}: graph_cover {t:9.1f} ms (call {wall * 1e3:9.1f} ms), {sw} sweeps, {cl.max()} clusters, {N * sw / t * 1e-3:.2f} M node visits/s, identical to first: {same}", flush=True)
