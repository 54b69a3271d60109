/*
 * mcb200_glue.c -- the `.Call` shim a metacell maintainer adds to bind libmcb200.so (include/mcb200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: it needs R's headers (Rinternals.h), which are absent here
 * and on the GPU box.  It is deliberately mechanical: every function unpacks SEXPs into
 * the plain pointers of the C ABI, calls one or a few mcb_* entry points, and repacks the result.  All
 * arithmetic stays behind the ABI.  Build inside the R package with
 *     PKG_LIBS = -L<dir> -lmcb200      (src/Makevars)
 *
 * Error convention of the reference: stop("MC-ERR: ...") (R/coclust_graph_cov.r:24, R/gset.r:157) ->
 * Rf_error("%s", mcb_last_error()); library messages already start with "MC-ERR:" where they restate a
 * reference check.  Indices cross the boundary 0-based and become 1-based here.
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <stdlib.h>
#include "mcb200.h"

/* One session per R process: a context + NCCL communicator + host thread for every device named by
 * options(mcb200.devices = c(0, 1, ...)) (default: device 0 alone).  The R process stays single-threaded; the
 * session's threads live inside the library and are joined before every call returns (SURVEY 8b "Threading"). */
static mcb_multi* g_multi = NULL;

static mcb_multi* session(void)
{
    if (!g_multi) {
        SEXP opt = GetOption1(install("mcb200.devices"));
        int32_t dev[64] = {0}, n = 1;
        if (!isNull(opt)) {
            SEXP iv = PROTECT(coerceVector(opt, INTSXP));
            n = 0;
            for (R_xlen_t k = 0; k < XLENGTH(iv) && n < 64; ++k) dev[n++] = INTEGER(iv)[k];
            UNPROTECT(1);
        }
        if (mcb_multi_create(dev, n, &g_multi) != 0) Rf_error("%s", mcb_last_error());
    }
    return g_multi;
}
static mcb_ctx* ctx(void) { return mcb_multi_ctx(session(), 0); }       /* single-device calls run on the first device */
#define CHECK(call) do { if ((call) != 0) Rf_error("%s", mcb_last_error()); } while (0)

/* called from .onUnload */
SEXP mcbr_shutdown(void) { mcb_multi_destroy(g_multi); g_multi = NULL; return R_NilValue; }

/* dgCMatrix@p is int32 in R; the ABI takes int64 column pointers */
static int64_t* widen_p(SEXP p, R_xlen_t* ncells)
{
    R_xlen_t n = XLENGTH(p);
    int64_t* out = (int64_t*)R_alloc((size_t)n, sizeof(int64_t));
    for (R_xlen_t k = 0; k < n; ++k) out[k] = INTEGER(p)[k];
    *ncells = n - 1;
    return out;
}

/* scm_downsamp(umis, n) (reference R/utils.r:30-69): returns list(x = numeric(nnz), kept = logical(ncells)) */
SEXP mcbr_downsample(SEXP p, SEXP i, SEXP x, SEXP ngenes, SEXP n, SEXP seed, SEXP add_non_dsamp)
{
    R_xlen_t ncells;
    int64_t* p64 = widen_p(p, &ncells);
    mcb_csc *m = NULL, *d = NULL;
    CHECK(mcb_csc_upload(ctx(), asInteger(ngenes), (int64_t)ncells, p64, INTEGER(i), REAL(x), &m));
    if (mcb_csc_downsample(ctx(), m, asReal(n), (uint64_t)asReal(seed), asLogical(add_non_dsamp), &d) != 0) {
        mcb_csc_free(m);
        Rf_error("%s", mcb_last_error());
    }
    SEXP out_x = PROTECT(allocVector(REALSXP, XLENGTH(x)));
    SEXP kept = PROTECT(allocVector(LGLSXP, ncells));
    uint8_t* k8 = (uint8_t*)R_alloc((size_t)ncells, 1);
    int rc = mcb_csc_download(ctx(), d, REAL(out_x), k8);
    mcb_csc_free(d);
    mcb_csc_free(m);
    if (rc != 0) Rf_error("%s", mcb_last_error());
    for (R_xlen_t c = 0; c < ncells; ++c) LOGICAL(kept)[c] = k8[c];
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"x", "kept", ""}));
    SET_VECTOR_ELT(res, 0, out_x);
    SET_VECTOR_ELT(res, 1, kept);
    UNPROTECT(3);
    return res;
}

/* colSums + scm_which_downsamp_n (reference R/scmat.r:399-407) */
SEXP mcbr_downsamp_depth(SEXP col_sums)
{
    R_xlen_t n = XLENGTH(col_sums);
    int64_t* cs = (int64_t*)R_alloc((size_t)n, sizeof(int64_t));
    for (R_xlen_t k = 0; k < n; ++k) cs[k] = (int64_t)REAL(col_sums)[k];
    double depth = 0;
    CHECK(mcb_downsamp_depth(cs, (int64_t)n, &depth));
    return ScalarReal(depth);
}

static SEXP edges_to_list(int64_t ne, const int32_t* src, const int32_t* dst, const double* w, int64_t nvalid)
{
    SEXP mc1 = PROTECT(allocVector(INTSXP, ne)), mc2 = PROTECT(allocVector(INTSXP, ne)), ww = PROTECT(allocVector(REALSXP, ne));
    for (int64_t e = 0; e < ne; ++e) { INTEGER(mc1)[e] = src[e] + 1; INTEGER(mc2)[e] = dst[e] + 1; REAL(ww)[e] = w[e]; }
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"mc1", "mc2", "w", "n_valid", ""}));
    SET_VECTOR_ELT(res, 0, mc1); SET_VECTOR_ELT(res, 1, mc2); SET_VECTOR_ELT(res, 2, ww);
    SET_VECTOR_ELT(res, 3, ScalarReal((double)nvalid));
    UNPROTECT(4);
    return res;
}

/* body of mcell_add_cgraph_from_mat_bknn (reference R/cgraph_knn.r:10-25): sparse matrix + feature rows in,
 * (mc1, mc2, w) integer codes into colnames(mat) out */
SEXP mcbr_cgraph_bknn(SEXP p, SEXP i, SEXP x, SEXP ngenes, SEXP feat_rows, SEXP K, SEXP k_expand, SEXP k_beta,
                      SEXP k_nonz, SEXP dsamp, SEXP seed)
{
    R_xlen_t ncells;
    int64_t* p64 = widen_p(p, &ncells);
    R_xlen_t G = XLENGTH(feat_rows);
    int32_t* fr = (int32_t*)R_alloc((size_t)G, sizeof(int32_t));
    for (R_xlen_t g = 0; g < G; ++g) fr[g] = INTEGER(feat_rows)[g] - 1;
    int64_t ne = 0, nvalid = 0;
    int32_t *src = NULL, *dst = NULL;
    double* w = NULL;
    /* every device of the session takes a cell range of the dgCMatrix and writes its own rows of the edge table */
    CHECK(mcb_cgraph_bknn_multi(session(), asInteger(ngenes), (int64_t)ncells, p64, INTEGER(i), REAL(x), fr, (int32_t)G, asInteger(K),
                                asInteger(k_expand), asInteger(k_beta), asReal(k_nonz), asLogical(dsamp), (uint64_t)asReal(seed),
                                &ne, &src, &dst, &w, &nvalid));
    SEXP res = edges_to_list(ne, src, dst, w, nvalid);
    mcb_free(src); mcb_free(dst); mcb_free(w);
    return res;
}

/* tgs_cor_graph(x, knn, k_expand, k_alpha, k_beta) (reference R/utils.r:115-123) on a dense genes x cells matrix */
SEXP mcbr_cor_graph_dense(SEXP x, SEXP K, SEXP k_expand, SEXP k_beta)
{
    SEXP dim = getAttrib(x, R_DimSymbol);
    const int32_t G = INTEGER(dim)[0];
    const int64_t ncells = INTEGER(dim)[1];
    mcb_cgraph* g = NULL;
    CHECK(mcb_cgraph_create_dense(ctx(), REAL(x), G, ncells, &g));
    int64_t ne = 0, nvalid = 0;
    int rc = mcb_cgraph_knn(g, asInteger(K), asInteger(k_expand), 0, 1);
    if (!rc) rc = mcb_cgraph_mutual(g, asInteger(k_beta));
    if (!rc) rc = mcb_cgraph_prune(g, &ne);
    if (rc) { mcb_cgraph_destroy(g); Rf_error("%s", mcb_last_error()); }
    int32_t* src = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* dst = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    double* w = (double*)R_alloc((size_t)(ne ? ne : 1), sizeof(double));
    rc = mcb_cgraph_edges(g, src, dst, w);
    mcb_cgraph_num_valid(g, &nvalid);
    mcb_cgraph_destroy(g);
    if (rc) Rf_error("%s", mcb_last_error());
    return edges_to_list(ne, src, dst, w, nvalid);
}

/* per-gene block of scm_gene_stat (reference R/gstats.r:95-131, :143-151): raw dgCMatrix slots, depth n, seed,
 * niche quantile, K_std_n and the (1-based) cells sz_cor is computed on (NULL = all).  Returns a 12 x ngenes matrix,
 * rows in the order documented at mcb_gene_stats (include/mcb200.h). */
SEXP mcbr_gene_stats(SEXP p, SEXP i, SEXP x, SEXP ngenes, SEXP n, SEXP seed, SEXP niche_quantile, SEXP k_std_n, SEXP sub_cells)
{
    R_xlen_t ncells;
    int64_t* p64 = widen_p(p, &ncells);
    mcb_csc *m = NULL, *ds = NULL;
    CHECK(mcb_csc_upload(ctx(), asInteger(ngenes), ncells, p64, INTEGER(i), REAL(x), &m));
    if (mcb_csc_downsample(ctx(), m, asReal(n), (uint64_t)asReal(seed), 0, &ds) != 0) { mcb_csc_free(m); Rf_error("%s", mcb_last_error()); }
    int32_t* sub = NULL; int64_t nsub = 0;
    if (!isNull(sub_cells)) {
        nsub = XLENGTH(sub_cells);
        sub = (int32_t*)R_alloc((size_t)(nsub ? nsub : 1), sizeof(int32_t));
        for (int64_t q = 0; q < nsub; ++q) sub[q] = INTEGER(sub_cells)[q] - 1;
    }
    SEXP out = PROTECT(allocMatrix(REALSXP, asInteger(ngenes), 12));          /* column-major: one column per statistic */
    int rc = mcb_gene_stats(ctx(), m, ds, asReal(niche_quantile), asReal(k_std_n), sub, nsub, REAL(out));
    mcb_csc_free(ds); mcb_csc_free(m);
    UNPROTECT(1);
    if (rc) Rf_error("%s", mcb_last_error());
    return out;
}

/* tgs_matrix_tapply(us, mc@mc, f) for f = geometric mean of 1+y minus 1 and f = mean(x > 0) (reference R/mc.r:185-186, :222, :240).
 * mc: integer metacell id per column of the matrix (1-based, NA for cells outside the cover).  Returns
 * list(geomean, frac_on) as ngenes x n_mc matrices, mc_size and mc_meansize. */
SEXP mcbr_mc_gene_tapply(SEXP p, SEXP i, SEXP x, SEXP ngenes, SEXP mc, SEXP n_mc)
{
    R_xlen_t ncells;
    int64_t* p64 = widen_p(p, &ncells);
    const int32_t ng = asInteger(ngenes), nm = asInteger(n_mc);
    int32_t* ids = (int32_t*)R_alloc((size_t)(ncells ? ncells : 1), sizeof(int32_t));
    for (R_xlen_t c = 0; c < ncells; ++c) ids[c] = INTEGER(mc)[c] == NA_INTEGER ? -1 : INTEGER(mc)[c] - 1;
    mcb_csc* m = NULL;
    CHECK(mcb_csc_upload(ctx(), ng, ncells, p64, INTEGER(i), REAL(x), &m));
    SEXP geo = PROTECT(allocMatrix(REALSXP, ng, nm)), frac = PROTECT(allocMatrix(REALSXP, ng, nm));     /* column-major == [n_mc][ngenes] */
    SEXP msz = PROTECT(allocVector(REALSXP, nm)), mmean = PROTECT(allocVector(REALSXP, nm));
    int64_t* sz = (int64_t*)R_alloc((size_t)nm, sizeof(int64_t));
    int rc = mcb_mc_gene_tapply(ctx(), m, ids, nm, REAL(geo), REAL(frac), sz, REAL(mmean));
    mcb_csc_free(m);
    if (rc) { UNPROTECT(4); Rf_error("%s", mcb_last_error()); }
    for (int32_t k = 0; k < nm; ++k) REAL(msz)[k] = (double)sz[k];
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"geomean", "frac_on", "mc_size", "mc_meansize", ""}));
    SET_VECTOR_ELT(res, 0, geo); SET_VECTOR_ELT(res, 1, frac); SET_VECTOR_ELT(res, 2, msz); SET_VECTOR_ELT(res, 3, mmean);
    UNPROTECT(5);
    return res;
}

/* tgs_cor_knn(x, y, knn) with y != x (reference R/atlas.r:164-166,277,475-478,584): x, y dense genes x cells */
SEXP mcbr_cor_knn_xy(SEXP x, SEXP y, SEXP knn)
{
    SEXP dx = getAttrib(x, R_DimSymbol), dy = getAttrib(y, R_DimSymbol);
    if (INTEGER(dx)[0] != INTEGER(dy)[0]) Rf_error("MC-ERR: tgs_cor_knn: x and y must have the same number of rows");
    const int32_t G = INTEGER(dx)[0], k = asInteger(knn);
    const int64_t nx = INTEGER(dx)[1], ny = INTEGER(dy)[1];
    SEXP col = PROTECT(allocMatrix(INTSXP, k, nx)), cor = PROTECT(allocMatrix(REALSXP, k, nx));   /* [nx][knn] row-major */
    float* cf = (float*)R_alloc((size_t)nx * k, sizeof(float));
    if (mcb_cor_knn_xy(ctx(), REAL(x), nx, REAL(y), ny, G, k, INTEGER(col), cf) != 0) { UNPROTECT(2); Rf_error("%s", mcb_last_error()); }
    for (int64_t q = 0; q < nx * k; ++q) {
        REAL(cor)[q] = INTEGER(col)[q] < 0 ? NA_REAL : (double)cf[q];
        INTEGER(col)[q] = INTEGER(col)[q] < 0 ? NA_INTEGER : INTEGER(col)[q] + 1;
    }
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"col", "cor", ""}));
    SET_VECTOR_ELT(res, 0, col); SET_VECTOR_ELT(res, 1, cor);
    UNPROTECT(3);
    return res;
}

/* tgs_graph_cover_resample(edges, knn, ..., method = "full") (reference R/coclust_graph_cov.r:41).
 * samples: integer matrix n_s x n_resamp (one resample per COLUMN, 1-based, ascending); seeds: numeric.
 * The resamples are strided over the session's devices and the counters reduced inside the library. */
SEXP mcbr_coclust(SEXP n_nodes, SEXP col1, SEXP col2, SEXP weight, SEXP knn, SEXP samples, SEXP seeds, SEXP min_mc_size,
                  SEXP cooling, SEXP burn_in)
{
    const int32_t N = asInteger(n_nodes);
    const int64_t ne = XLENGTH(col1);
    SEXP sdim = getAttrib(samples, R_DimSymbol);
    const int32_t n_s = INTEGER(sdim)[0], n_resamp = INTEGER(sdim)[1];
    int32_t* src = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* dst = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    for (int64_t e = 0; e < ne; ++e) { src[e] = INTEGER(col1)[e] - 1; dst[e] = INTEGER(col2)[e] - 1; }
    int32_t* smp = (int32_t*)R_alloc((size_t)n_s * n_resamp, sizeof(int32_t));
    for (int64_t q = 0; q < (int64_t)n_s * n_resamp; ++q) smp[q] = INTEGER(samples)[q] - 1;     /* column-major == [n_resamp][n_s] */
    uint64_t* sd = (uint64_t*)R_alloc((size_t)n_resamp, sizeof(uint64_t));
    for (int32_t r = 0; r < n_resamp; ++r) sd[r] = (uint64_t)REAL(seeds)[r];
    int64_t np = 0;
    int32_t *h1 = NULL, *h2 = NULL, *hc = NULL, *hn = NULL;
    CHECK(mcb_coclust_multi(session(), N, ne, src, dst, REAL(weight), n_resamp, n_s, smp, sd, asInteger(min_mc_size), asReal(cooling),
                            asInteger(burn_in), 1000, asInteger(knn), &np, &h1, &h2, &hc, &hn));
    SEXP n1 = PROTECT(allocVector(INTSXP, np)), n2 = PROTECT(allocVector(INTSXP, np)), cnt = PROTECT(allocVector(INTSXP, np));
    SEXP ns = PROTECT(allocVector(INTSXP, N));
    for (int64_t q = 0; q < np; ++q) { INTEGER(n1)[q] = h1[q] + 1; INTEGER(n2)[q] = h2[q] + 1; INTEGER(cnt)[q] = hc[q]; }
    for (int32_t v = 0; v < N; ++v) INTEGER(ns)[v] = hn[v];
    mcb_free(h1); mcb_free(h2); mcb_free(hc); mcb_free(hn);
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"node1", "node2", "cnt", "samples", ""}));
    SET_VECTOR_ELT(res, 0, n1); SET_VECTOR_ELT(res, 1, n2); SET_VECTOR_ELT(res, 2, cnt); SET_VECTOR_ELT(res, 3, ns);
    UNPROTECT(5);
    return res;
}

/* method = "edges" of tgs_graph_cover_resample: co-occurrence of graph edges only, O(edges) memory (for cell counts whose
 * N x N matrix cannot exist); rows with cnt == 0 are dropped here.  Single device. */
SEXP mcbr_coclust_edges(SEXP n_nodes, SEXP col1, SEXP col2, SEXP weight, SEXP knn, SEXP samples, SEXP seeds, SEXP min_mc_size,
                        SEXP cooling, SEXP burn_in)
{
    const int32_t N = asInteger(n_nodes);
    const int64_t ne = XLENGTH(col1);
    SEXP sdim = getAttrib(samples, R_DimSymbol);
    const int32_t n_s = INTEGER(sdim)[0], n_resamp = INTEGER(sdim)[1];
    int32_t* src = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* dst = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    for (int64_t e = 0; e < ne; ++e) { src[e] = INTEGER(col1)[e] - 1; dst[e] = INTEGER(col2)[e] - 1; }
    int32_t* smp = (int32_t*)R_alloc((size_t)n_s * n_resamp, sizeof(int32_t));
    for (int64_t q = 0; q < (int64_t)n_s * n_resamp; ++q) smp[q] = INTEGER(samples)[q] - 1;
    uint64_t* sd = (uint64_t*)R_alloc((size_t)n_resamp, sizeof(uint64_t));
    for (int32_t r = 0; r < n_resamp; ++r) sd[r] = (uint64_t)REAL(seeds)[r];
    mcb_coclust* c = NULL;
    CHECK(mcb_coclust_run_edges(ctx(), N, ne, src, dst, REAL(weight), n_resamp, n_s, smp, sd, asInteger(min_mc_size), asReal(cooling),
                                asInteger(burn_in), 1000, asInteger(knn), 0, 1, &c));
    int32_t* e1 = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* e2 = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* ec = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    SEXP ns = PROTECT(allocVector(INTSXP, N));
    const int rc = mcb_coclust_edge_counts(c, e1, e2, ec, INTEGER(ns));
    mcb_coclust_destroy(c);
    if (rc) { UNPROTECT(1); Rf_error("%s", mcb_last_error()); }
    int64_t np = 0;
    for (int64_t e = 0; e < ne; ++e) np += ec[e] > 0;
    SEXP n1 = PROTECT(allocVector(INTSXP, np)), n2 = PROTECT(allocVector(INTSXP, np)), cnt = PROTECT(allocVector(INTSXP, np));
    for (int64_t e = 0, q = 0; e < ne; ++e)
        if (ec[e] > 0) { INTEGER(n1)[q] = e1[e] + 1; INTEGER(n2)[q] = e2[e] + 1; INTEGER(cnt)[q] = ec[e]; ++q; }
    SEXP res = PROTECT(mkNamed(VECSXP, (const char*[]){"node1", "node2", "cnt", "samples", ""}));
    SET_VECTOR_ELT(res, 0, n1); SET_VECTOR_ELT(res, 1, n2); SET_VECTOR_ELT(res, 2, cnt); SET_VECTOR_ELT(res, 3, ns);
    UNPROTECT(5);
    return res;
}

/* tgs_graph_cover(edges, min_mc_size, cooling, burn_in) (reference R/mc_graphcov.r:27, R/mc_from_coclust.r:247):
 * one cover over all nodes; returns integer cluster per node, 0 = outlier (R/mc_graphcov.r:28) */
SEXP mcbr_graph_cover(SEXP n_nodes, SEXP col1, SEXP col2, SEXP weight, SEXP min_mc_size, SEXP cooling, SEXP burn_in, SEXP seed)
{
    const int32_t N = asInteger(n_nodes);
    const int64_t ne = XLENGTH(col1);
    int32_t* src = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    int32_t* dst = (int32_t*)R_alloc((size_t)(ne ? ne : 1), sizeof(int32_t));
    for (int64_t e = 0; e < ne; ++e) { src[e] = INTEGER(col1)[e] - 1; dst[e] = INTEGER(col2)[e] - 1; }
    SEXP cl = PROTECT(allocVector(INTSXP, N));
    int32_t sweeps = 0;
    if (mcb_graph_cover(ctx(), N, ne, src, dst, REAL(weight), asInteger(min_mc_size), asReal(cooling), asInteger(burn_in), 1000,
                        (uint64_t)asReal(seed), INTEGER(cl), &sweeps) != 0) { UNPROTECT(1); Rf_error("%s", mcb_last_error()); }
    UNPROTECT(1);
    return cl;
}
