#!/usr/bin/env python
"""bench.py -- headline benchmark of the balanced K-nn cell-graph path on B200.

Metric (BASELINE.json): cells/sec for the balanced K=100 cgraph build.  One "step" is one full pass
of the hot path (features -> correlation + top-Kc -> rank balancing -> edge table) over the C2
workload: 100,000 cells x 1,000 feature genes, K = 100 (k_expand 10, k_beta 3), synthetic
negative-binomial UMIs.  At N > 1 the same workload is sharded over N ranks (strong scaling): every rank
contracts 1/N of the symmetric tile list, candidate records reach their row owners in one NCCL all-to-all, and
the kNN lists / in-degree cuts are all-gathered.

  value : cells / s with the CSC matrix already resident in HBM (device-timed, max over ranks)
  e2e   : cells / s through the C ABI with host (pinned) buffers: H2D of the CSC + the whole path +
          D2H of the (mc1, mc2, w) edge table inside the timed region
  roofline     : the dominant kernel (tcgen05 correlation + fused filter), CUDA-event timed
  cpu_baseline : CPU port of the dominant stage (OpenBLAS fp64 dgemm + partial selection) on a bounded
                 row sample, all host cores
  coclust      : second metric, co-clustering resamples / s on the C3 workload (N = 1 only)

`--impl reference` times the CPU port alone (the reference's tgstat cannot run here: no R).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv or os.environ.get("MCB_BENCH_ALL_CORES"):
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all host cores, and the
    # BLAS / OpenMP pools read these when numpy and the oracle are first loaded
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec for balanced K=100 cgraph build"
UNIT = "cells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=1000)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--no-coclust", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check against the fp64 oracle")
    ap.add_argument("--parity-rows", type=int, default=256)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto ~15 s)")
    return ap.parse_args()


def workload_name(a):
    return f"C2-shaped synthetic NB UMIs: {a.cells} cells x {a.genes} feature genes, K={a.K}, k_expand=10, k_beta=3"


def make_input(a):
    from metacell_b200 import synth
    m = synth.make_umi_matrix(a.cells, a.genes, n_types=24, median_umis=1500.0, seed=a.seed)
    return m, np.arange(a.genes, dtype=np.int32)


# ------------------------------------------------------------------------------------------------
# CPU port (baseline + reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_port(a, m, feats, rows=0, target_s=15.0):
    """A3+A4 on all cells, then A5 (the >95% stage) on a bounded row sample with all host cores: OpenBLAS fp64 dgemm in
    row blocks + OpenMP per-row partial selection (oracle/mc_oracle.c orc_select_rows)."""
    from oracle import oracle as O, oracle_np as P
    fm = np.full(m.ngenes, -1, np.int32)
    fm[feats] = np.arange(feats.size)
    t0 = time.perf_counter()
    Z, valid = O.features(m.indptr, m.indices, m.data, fm, feats.size)
    Zv = Z[valid]
    t_feat = time.perf_counter() - t0
    M = Zv.shape[0]
    Kc = a.K * 10
    if rows <= 0:           # calibrate: one block, then size the sample for ~15 s
        t0 = time.perf_counter()
        P.cor_knn_blas(Zv, Kc, 0, min(1024, M))
        per_row = (time.perf_counter() - t0) / min(1024, M)
        rows = int(max(1024, min(M, target_s / max(per_row, 1e-9)) // 1024 * 1024))
    rows = min(rows, M)
    split = {}
    t0 = time.perf_counter()
    P.cor_knn_blas(Zv, Kc, 0, rows, timing=split)
    t_knn = time.perf_counter() - t0
    cells_per_s = rows / (t_knn + t_feat * rows / M)
    return dict(value=cells_per_s, unit=UNIT, cores=os.cpu_count(), omp_threads=O.num_threads(), kind="port",
                phases_s={"features_all_cells": round(t_feat, 3), "dgemm": round(split.get("dgemm", 0.0), 3),
                          "select": round(split.get("select", 0.0), 3)},
                sample=f"features on all {M} cells + correlation/top-{Kc} (OpenBLAS fp64 dgemm + OpenMP per-row select) for "
                       f"{rows} of {M} rows against all columns; balancing (<3% of CPU time) excluded; "
                       f"{t_knn:.1f} s measured"), rows, t_knn


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, feats = make_input(a)
    vals = []
    rows = a.cpu_rows
    target = min(15.0, 150.0 / max(1, a.warmup + a.steps))      # the whole run ends within a few minutes
    for it in range(a.warmup + a.steps):
        cb, rows, t = cpu_port(a, m, feats, rows, target)
        if it >= a.warmup:
            vals.append((cb["value"], t))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals]) * 1e3)
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(a)},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU port of the tgstat path (tgstat itself cannot run here: no R); each step is a bounded row sample"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# in-bench parity: the bench workload's own result against the fp64 oracle (untimed, after the timed region)
# ------------------------------------------------------------------------------------------------
def parity_block(a, m, feats, cells, col, cor, src, dst, w, row0=0):
    """`col` / `cor`: the GPU's kNN lists of rows [row0, row0 + col.shape[0]) of the valid cells; (src, dst, w) the
    balanced edges of the same rows.  Checks a random row sample against orc_select_rows on fp64 correlations
    (north_star: correlations within 1e-5, lists equal except inside runs tied within that tolerance) and, when the
    lists cover every row, the balanced edges bit for bit against orc_balance of the GPU's own lists."""
    from oracle import oracle as O, oracle_np as P
    t0 = time.perf_counter()
    fm = np.full(m.ngenes, -1, np.int32)
    fm[feats] = np.arange(feats.size)
    Z, valid = O.features(m.indptr, m.indices, m.data, fm, feats.size)
    Zv = Z[valid]
    M, Kc = Zv.shape[0], a.K * 10
    ok_cells = bool(np.array_equal(cells, np.nonzero(valid)[0]))
    n = min(a.parity_rows, col.shape[0])
    pick = np.sort(np.random.default_rng(a.seed + 99).choice(col.shape[0], n, replace=False))
    ocol, ocor = P.cor_knn_blas(Zv, Kc + 8, rows=pick + row0)
    gc, gr = col[pick], cor[pick].astype(np.float64)
    dcor = float(np.abs(gr - ocor[:, :Kc]).max())
    diff = gc != ocol[:, :Kc]
    off_tie = 0                       # disagreeing positions whose oracle neighbours are NOT tied within 2e-5
    for i, j in zip(*np.nonzero(diff)):
        nb = [abs(ocor[i, j] - ocor[i, jj]) for jj in (j - 1, j + 1) if 0 < jj < ocor.shape[1]]
        if not nb or min(nb) >= 2e-5:
            off_tie += 1
    out = {"rows": int(n), "of_rows": int(M), "max_abs_dcor": dcor, "list_agreement": float((~diff).mean()),
           "disagreements_off_ties": int(off_tie), "valid_cells_equal": ok_cells, "tolerance": 1e-5}
    if col.shape[0] == M:
        deg, odst, _ = O.balance(col, a.K, 10, 3)
        osrc, odst2, ow = O.edges_from_balance(deg, odst, a.K)
        out["edges_bit_exact_on_gpu_lists"] = bool(np.array_equal(src, cells[osrc]) and np.array_equal(dst, cells[odst2])
                                                   and np.array_equal(w, ow))
        out["edges"] = int(src.size)
    out["ok"] = bool(ok_cells and dcor < 1e-5 and off_tie == 0 and out.get("edges_bit_exact_on_gpu_lists", True))
    out["seconds"] = round(time.perf_counter() - t0, 1)
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  In-process NVML queries from a background
    thread (an `nvidia-smi -lms` loop was measured to stall the driver and triple the step time)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, dev, period=float(os.environ.get("MCB_BENCH_CLOCK_PERIOD", "0.05"))):
        self.dev, self.period, self.sm, self.bits, self.mx, self.ok = dev, period, [], 0, None, False
        self.stop_flag = threading.Event()

    def start(self):
        if os.environ.get("MCB_BENCH_NO_SMI"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.dev)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:           # noqa: BLE001
            self.err = str(e)

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:            # noqa: BLE001
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.th.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "samples": len(self.sm),
                "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def _ncu_traffic(a, M):
    """dram bytes read + written by one launch of the dominant kernel, from the newest committed `ncu --set full`
    summary under profiles/ (captured on this same workload; None when the run is a different shape or N > 1)."""
    import glob
    if a.cells != 100000 or a.genes != 1000 or a.gpus != 1:
        return None
    best = None
    for fn in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_ncu_tc_cor_kernel.csv"))):
        tot, name = 0.0, ""
        for ln in open(fn):
            f = [q.strip('"') for q in ln.rstrip("\n").split('","')]
            if len(f) >= 3 and f[0].strip('"') in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2].strip('"')) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(f[1], 1.0)
            if len(f) >= 3 and f[0].strip('"') == "Kernel Name":
                name = f[2]
        if tot > 0 and "tc_cor2_kernel" in name:
            best = {"bytes": tot, "source": os.path.basename(fn)}
    return best


def run_ours(a):
    import torch
    import torch.distributed as dist
    from metacell_b200 import _lib
    from metacell_b200 import dist as mdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)

    m, feats = make_input(a)
    nnz = m.nnz
    # pinned host copies of the dgCMatrix slots (what R would hand over) + device-resident copy
    hp = torch.from_numpy(m.indptr.astype(np.int64)).pin_memory()
    hi = torch.from_numpy(m.indices.astype(np.int32)).pin_memory()
    hx = torch.from_numpy(m.data.astype(np.float64)).pin_memory()
    dp, di = hp.cuda(), hi.cuda()
    dx = torch.from_numpy(m.data.astype(np.int64)).cuda().to(torch.int32)      # uint32 counts on device
    csc_dev = _lib.Csc.wrap_device(ctx, m.ngenes, m.ncells, nnz, dp.data_ptr(), di.data_ptr(), dx.data_ptr(), keep=(dp, di, dx))
    max_edges = m.ncells * a.K
    out_src = torch.empty(max_edges, dtype=torch.int32).pin_memory()
    out_dst = torch.empty(max_edges, dtype=torch.int32).pin_memory()
    out_w = torch.empty(max_edges, dtype=torch.float64).pin_memory()
    out_np = (out_src.numpy(), out_dst.numpy(), out_w.numpy())

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    info = {}

    def step(csc, d2h):
        t_dbg = time.perf_counter()
        g = _lib.CGraph.from_csc(ctx, csc, feats, 7.0)
        eng = mdist.CudaCGraphEngine(g)
        mdist.cgraph_sharded(eng, a.K, 10, 3, rank, world, fetch_edges=False)
        if d2h:
            g.edges(out_np)
        info["edges"] = g.n_edges
        info["stats"] = g.knn_stats()
        info["rows"] = g.rows
        info["M"] = g.M
        g.destroy()
        if os.environ.get("MCB_BENCH_DEBUG"):
            print(f"[debug] step wall {1e3 * (time.perf_counter() - t_dbg):.1f} ms", file=sys.stderr, flush=True)

    def timed(n_steps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record(stream)
        for _ in range(n_steps):
            if e2e:
                csc = _lib.Csc.upload_ptrs(ctx, m.ngenes, m.ncells, nnz, hp.data_ptr(), hi.data_ptr(), hx.data_ptr())
                step(csc, True)
                csc.free()
            else:
                step(csc_dev, False)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.launch_count() - l0

    for _ in range(a.warmup):
        step(csc_dev, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(a.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    # per-stage CUDA-event timings of one more (untimed) step, for the roofline of the dominant kernel
    ctx.set_option("profile", 1)
    ctx.timings_reset()
    step(csc_dev, False)
    stages = ctx.timings()
    ctx.set_option("profile", 0)
    # end to end through host buffers
    timed(1, True)
    ms_e2e, _ = timed(max(1, a.steps), True)
    n_e2e = max(1, a.steps)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ms_step = ms_total / a.steps
        value = a.cells / (ms_step * 1e-3)
        e2e_v = a.cells / (ms_e2e / n_e2e * 1e-3)
        gemm_ms, _ = stages.get("cor_filter_gemm", (float("nan"), 0))
        M, rows = info["M"], info["rows"]
        flop = 2.0 * rows * M * a.genes                      # algorithmic: fp32-equivalent, unpadded G, this rank's rows
        achieved = flop / (gemm_ms * 1e-3) / 1e12
        bf16 = peaks.get("bf16_tflops_sustained")
        peak = bf16 if bf16 else 1400.0
        roof = {"bound": "tensor", "kernel": "tc_cor2_kernel<SYM> (tcgen05 cta_group::2 kind::f16, 3-pass hi/lo split of z*2^12, fp32 accumulate, fused top-Kc filter)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": (_ncu_traffic(a, M) or {}).get("bytes"), "traffic_unit": "bytes of DRAM read + written per launch (ncu --set full)",
                "traffic_source": (_ncu_traffic(a, M) or {}).get("source"),
                "issued_tflops": 3.0 * achieved * (-(-a.genes // 64) * 64) / a.genes * 0.5 * (1.0 + 256.0 / max(M, 256)),
                "ms_per_launch": gemm_ms, "share_of_step": gemm_ms / ms_step,
                "peak_source": ("bf16_tflops_sustained of MEASURED_PEAKS.json (kind::f16 runs at the bf16 rate; kernel timed inside a long step)"
                                if bf16 else "1400 TFLOP/s sustained fallback (B200_PROFILING.md)"),
                "note": "algorithmic flop = 2*rows*M*G (full square, fp32-equivalent); 3 passes are issued per pair and the symmetric schedule contracts each unordered pair once (across all ranks), so issued = 1.5x algorithmic"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16x3 split (fp32-equivalent, fp32 accumulate); int32/uint64 index work", "data": "synthetic",
                "config": {"workload": workload_name(a), "l2": "inputs larger than L2 (Z planes 0.8 GB, candidate buffers 1.6 GB)",
                           "parallelism": f"symmetric tile list and rows sharded over {world} rank(s); exchanges: thresholds + kNN lists + in-degree cuts (all-gather), candidate records (all-to-all)", "valid_cells": M, "edges": info["edges"],
                           "knn_filter": info["stats"]},
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(hp.numel() * 8 + hi.numel() * 4 + hx.numel() * 8),
                        "d2h_bytes_per_step": int(info["edges"] * 16), "ms_per_step": ms_e2e / n_e2e},
                "gpu_launches": launches, "clocks": clocks,
                "stages_ms": {k: round(v[0], 3) for k, v in stages.items()}, "roofline": roof}
        if world == 1 and not a.no_cpu:
            cb, _, _ = cpu_port(a, m, feats, a.cpu_rows)
            line["cpu_baseline"] = cb
        if world == 1 and not a.no_parity:
            g = _lib.CGraph.from_csc(ctx, csc_dev, feats, 7.0)
            cells = g.valid_cells()
            g.knn(a.K, 10)
            col, cor = g.download_knn()
            g.mutual(3)
            g.prune()
            src, dst, w = g.edges()
            g.destroy()
            line["parity"] = parity_block(a, m, feats, cells, col, cor, src, dst, w)
            del col, cor
    coc = None if a.no_coclust else bench_coclust(ctx, a, rank, world)
    if rank == 0:
        if coc is not None:
            line["coclust"] = coc
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_coclust(ctx, a, rank=0, world=1):
    """C3: PBMC8k-shaped graph (8,000 cells, K=100), 500 resamples at p=0.75, min_mc_size=20.  At world > 1 the
    resamples are strided over the ranks and the counters are reduced to rank 0 over NCCL."""
    import torch
    import torch.distributed as dist
    from metacell_b200 import _lib, synth
    from metacell_b200 import dist as mdist
    m = synth.make_umi_matrix(8000, 600, n_types=12, median_umis=3000.0, seed=a.seed + 5)
    csc = _lib.Csc.upload(ctx, m.ngenes, m.indptr, m.indices, m.data.astype(np.float64))
    g = _lib.CGraph.from_csc(ctx, csc, np.arange(600, dtype=np.int32), 7.0)
    g.knn(100, 10); g.mutual(3); g.prune()
    src, dst, w = g.edges()
    g.destroy(); csc.free()
    sets, seeds = synth.resample_sets(m.ncells, 0.75, 500, 77)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=ctx.device)

    def run():
        cc = _lib.CoClust(ctx, m.ncells, src, dst, w, sets, seeds, 20, 1.05, 10, rank=rank, nranks=world)
        out = mdist.coclust_sharded(cc, rank, world)
        return cc, out

    warm, _ = run()                      # warm-up at the timed shape (block cache, NCCL buffers)
    warm.destroy()
    if rank == 0:
        ctx.set_option("profile", 1); ctx.timings_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    e0.record(stream)
    cc, out = run()
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res = None
    if rank == 0:
        st = ctx.timings(); ctx.set_option("profile", 0)
        _, sweeps = cc.assignments()
        res = {"metric": "coclust resamples/sec", "value": 500.0 / (ms * 1e-3), "unit": "resamples/s", "ms": ms, "n_gpus": world,
               "config": {"workload": "C3: 8,000 cells, balanced K=100 graph, 500 resamples, p=0.75, min_mc_size=20",
                          "edges": int(src.size), "pairs": int(out[0].size), "mean_sweeps_rank0": float(np.mean(sweeps))},
               "stages_ms_rank0": {k: round(v[0], 3) for k, v in st.items()},
               "note": "host edge table in, host (node1,node2,cnt) out on rank 0; graph cover is latency-bound (DESIGN.md)"}
    cc.destroy()
    return res


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
