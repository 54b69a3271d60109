"""Synthetic negative-binomial UMI matrices of the shapes BASELINE.json names.

Host-side numpy only; used by tests/ and bench.py to build inputs.  Genes x cells, CSC, like the
`dgCMatrix` in `tgScMat@mat` (reference R/scmat.r).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class UmiMatrix:
    """CSC genes x cells count matrix (the three dgCMatrix slots + dims)."""
    indptr: np.ndarray      # int64 [ncells + 1]   (dgCMatrix@p)
    indices: np.ndarray     # int32 [nnz] gene row (dgCMatrix@i)
    data: np.ndarray        # uint32 [nnz] counts  (dgCMatrix@x, integral)
    ngenes: int
    ncells: int
    cell_type: np.ndarray   # int32 [ncells] latent type (ground truth for sanity checks)

    @property
    def nnz(self):
        return int(self.indptr[-1])

    def col_sums(self):
        cs = np.zeros(self.ncells, dtype=np.int64)
        np.add.at(cs, np.repeat(np.arange(self.ncells), np.diff(self.indptr)), self.data)
        return cs

    def to_dense(self):
        d = np.zeros((self.ngenes, self.ncells), dtype=np.uint32)
        cols = np.repeat(np.arange(self.ncells), np.diff(self.indptr))
        d[self.indices, cols] = self.data
        return d


def make_umi_matrix(ncells, ngenes, n_types=20, median_umis=3000.0, size_sigma=0.4, dispersion=3.0,
                    marker_frac=0.05, rare_types=0, seed=0, chunk=8192, cell_seed=None):
    """NB(mean = s_c * mu_g * fold[type_c, g], dispersion r) counts.

    mu_g log-normal, s_c log-normal(sigma=size_sigma); every type up-regulates a random
    `marker_frac` of genes 4-16x.  `rare_types` of the types get <0.5% frequency (C5 shape).
    `cell_seed` (optional) re-seeds the per-cell draws only, so several processes can generate different cell
    shards of one population (same gene means / type programs from `seed`).
    """
    rng = np.random.default_rng(seed)
    mu = rng.lognormal(mean=0.0, sigma=1.2, size=ngenes)
    fold = np.ones((n_types, ngenes), dtype=np.float64)
    nmark = max(1, int(marker_frac * ngenes))
    for t in range(n_types):
        g = rng.choice(ngenes, size=nmark, replace=False)
        fold[t, g] = 2.0 ** rng.uniform(2.0, 4.0, size=nmark)
    freq = rng.dirichlet(np.full(n_types, 3.0))
    if cell_seed is not None:
        rng = np.random.default_rng([seed, cell_seed])
    if rare_types > 0:
        freq[:rare_types] = rng.uniform(0.001, 0.005, size=rare_types)
        freq[rare_types:] *= (1.0 - freq[:rare_types].sum()) / freq[rare_types:].sum()
    ctype = rng.choice(n_types, size=ncells, p=freq / freq.sum()).astype(np.int32)
    base = (mu[None, :] * fold).astype(np.float64)            # [type, gene]
    base *= median_umis / np.median(base.sum(axis=1))
    size = rng.lognormal(mean=0.0, sigma=size_sigma, size=ncells)

    indptr = np.zeros(ncells + 1, dtype=np.int64)
    idx_parts, dat_parts = [], []
    for c0 in range(0, ncells, chunk):
        c1 = min(ncells, c0 + chunk)
        mean = base[ctype[c0:c1]] * size[c0:c1, None]          # [cells, genes]
        lam = rng.gamma(shape=dispersion, scale=mean / dispersion)
        cnt = rng.poisson(lam).astype(np.uint32)
        nz_c, nz_g = np.nonzero(cnt)
        indptr[c0 + 1:c1 + 1] = np.bincount(nz_c, minlength=c1 - c0)
        idx_parts.append(nz_g.astype(np.int32))
        dat_parts.append(cnt[nz_c, nz_g])
    indptr = np.cumsum(indptr)
    return UmiMatrix(indptr=indptr, indices=np.concatenate(idx_parts), data=np.concatenate(dat_parts),
                     ngenes=ngenes, ncells=ncells, cell_type=ctype)


def config(name, seed=1234, scale=1.0):
    """The BASELINE.json configs (C1..C5).  Returns (UmiMatrix, feature_genes, params)."""
    name = name.upper()
    if name == "C1":      # PBMC8k-shaped: 8,000 cells, ~600 feature genes among many rows, K=100, dsamp
        n, g = int(8000 * scale), 4000
        m = make_umi_matrix(n, g, n_types=12, median_umis=4500.0, seed=seed)
        feats = np.sort(np.random.default_rng(seed + 1).choice(g, size=600, replace=False)).astype(np.int32)
        return m, feats, dict(K=100, dsamp=True)
    if name == "C2":      # 100K cells x 1,000 feature genes (feature rows only), K=100
        n = int(100000 * scale)
        m = make_umi_matrix(n, 1000, n_types=24, median_umis=1500.0, seed=seed)
        return m, np.arange(1000, dtype=np.int32), dict(K=100, dsamp=False)
    if name == "C4":      # 1M cells x 1,000 feature genes, K=150
        n = int(1000000 * scale)
        m = make_umi_matrix(n, 1000, n_types=40, median_umis=1500.0, seed=seed)
        return m, np.arange(1000, dtype=np.int32), dict(K=150, dsamp=False)
    if name == "C5":      # Amphimedon-like: ~20K cells, low UMI/cell, many rare types, K=150
        n, g = int(20000 * scale), 3000
        m = make_umi_matrix(n, g, n_types=40, median_umis=800.0, size_sigma=0.6, rare_types=8, seed=seed)
        feats = np.sort(np.random.default_rng(seed + 1).choice(g, size=800, replace=False)).astype(np.int32)
        return m, feats, dict(K=150, dsamp=False)
    raise ValueError(f"unknown config {name}")


def resample_sets(ncells, p_resamp, n_resamp, seed):
    """Host-supplied node samples + per-resample seeds (north_star: 'host-supplied RNG seeds and
    resample index sets').  Sample size = round(p * N) (SURVEY App. A7 `resamp_size_rule`)."""
    rng = np.random.default_rng(seed)
    n_s = int(round(p_resamp * ncells))
    sets = np.empty((n_resamp, n_s), dtype=np.int32)
    for r in range(n_resamp):
        sets[r] = np.sort(rng.choice(ncells, size=n_s, replace=False))
    seeds = rng.integers(0, 2 ** 63 - 1, size=n_resamp, dtype=np.int64).astype(np.uint64)
    return sets, seeds
