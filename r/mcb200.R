# mcb200.R -- R side of the drop-in: same names and signatures as the reference functions they replace.
# NOT RUN IN THIS REPOSITORY (no R in the image, SURVEY.md 0.4).  Everything outside the arithmetic --
# scdb look-ups, get_param, messages, seed save/restore, object construction -- is unchanged reference
# behaviour; only the marked calls move behind .Call.

# replaces scm_downsamp (R/utils.r:30-69)
scm_downsamp = function(umis, n)
{
	old_seed = .set_seed(get_param("mc_rseed"))
	r = .Call("mcbr_downsample", umis@p, umis@i, umis@x, nrow(umis), as.numeric(n),
	          as.numeric(get_param("mc_rseed")), FALSE, PACKAGE = "mcb200")
	ds = umis
	ds@x = r$x
	ds = Matrix::drop0(ds[, r$kept])          # cells with < n UMIs are dropped (R/utils.r:34)
	.restore_seed(old_seed)
	return(ds)
}

# replaces tgs_cor_graph (R/utils.r:115-123); k_alpha stays accepted and ignored as in the reference
tgs_cor_graph = function(x, knn, k_expand, k_alpha, k_beta)
{
	r = .Call("mcbr_cor_graph_dense", x, as.integer(knn), as.integer(k_expand), as.integer(k_beta), PACKAGE = "mcb200")
	nms = colnames(x)
	data.frame(col1 = factor(nms[r$mc1], levels = nms), col2 = factor(nms[r$mc2], levels = nms), weight = r$w)
}

# replaces the stat_on_genes / doMC block of scm_gene_stat (R/gstats.r:95-151): one call returns the per-gene
# columns for every gene; the caller keeps f_g = tot > 10, the rounding, and the rollmedian trends (:153-190) in R
mcb_gene_stat_block = function(mat, downsample_n, niche_quantile = 0.2, K_std_n = 1000, subs = NULL)
{
	r = .Call("mcbr_gene_stats", mat@p, mat@i, mat@x, nrow(mat), as.numeric(downsample_n), as.numeric(get_param("mc_rseed")),
	          as.numeric(niche_quantile), as.numeric(K_std_n), if (is.null(subs)) NULL else as.integer(subs), PACKAGE = "mcb200")
	colnames(r) = c("tot", "var", "niche_stat", "n_mean", "is_on_count", "sz_cor", "ds_top1", "ds_top2", "ds_top3",
	                "ds_mean", "ds_var", "ds_is_on_count")
	data.frame(name = rownames(mat), r, stringsAsFactors = F)
}

# the two tgs_matrix_tapply reductions of mc_compute_fp / mc_compute_e_gc / mc_compute_cov_gc (R/mc.r:185-186, :222, :240)
# in one pass; e.g. in mc_compute_e_gc:  e_gc = mcb_mc_tapply(us, mc@mc)$geomean[f_g_cov, ]
mcb_mc_tapply = function(us, mc_of_cell)
{
	ids = as.integer(mc_of_cell[colnames(us)])
	r = .Call("mcbr_mc_gene_tapply", us@p, us@i, us@x, nrow(us), ids, max(ids, na.rm = T), PACKAGE = "mcb200")
	rownames(r$geomean) = rownames(us); rownames(r$frac_on) = rownames(us)
	r
}

# tgs_cor_knn for the y != x call sites (atlas projection, R/atlas.r:164-166,277,475-478,584); x == y keeps
# going through tgs_cor_graph / mcell_add_cgraph_from_mat_bknn.  Same data.frame schema as tgstat's.
tgs_cor_knn = function(x, y, knn, ...)
{
	r = .Call("mcbr_cor_knn_xy", x, y, as.integer(knn), PACKAGE = "mcb200")
	ok = !is.na(as.vector(r$col))
	nx = colnames(x); ny = colnames(y)
	data.frame(col1 = factor(nx[rep(1:ncol(x), each = knn)], levels = nx)[ok],
	           col2 = factor(ny[as.vector(r$col)], levels = ny)[ok],
	           cor = as.vector(r$cor)[ok], rank = rep(1:knn, ncol(x))[ok])
}

# replaces mcell_add_cgraph_from_mat_bknn (R/cgraph_knn.r:10-25): the sparse matrix and the feature rows go
# straight to the GPU, so the dense genes x cells double matrix (8 GB at 1M cells) is never built on the host
mcell_add_cgraph_from_mat_bknn = function(mat_id, gset_id, graph_id, K, dsamp = F)
{
	gset = scdb_gset(gset_id)
	if(is.null(gset)) stop("MC-ERR: non existing gset ", gset_id, " when trying to get feature matrix")
	mat = scdb_mat(mat_id)
	if(is.null(mat)) stop("MC-ERR: non existing mat ", mat_id, " when trying to get feature matrix")
	umis = mat@mat
	feat_rows = match(intersect(names(gset@gene_set), rownames(umis)), rownames(umis))     # R/gset.r:192-197
	message("will build balanced knn graph on ", ncol(umis), " cells and ", length(feat_rows),
	        " genes, this can be a bit heavy for >20,000 cells")
	r = .Call("mcbr_cgraph_bknn", umis@p, umis@i, umis@x, nrow(umis), as.integer(feat_rows), as.integer(K),
	          as.integer(get_param("scm_balance_graph_k_expand")), as.integer(get_param("scm_balance_graph_k_beta")),
	          as.numeric(get_param("scm_k_nonz_exp")), as.logical(dsamp), as.numeric(get_param("mc_rseed")),
	          PACKAGE = "mcb200")
	cnames = colnames(umis)
	gr = data.frame(mc1 = factor(cnames[r$mc1], levels = cnames), mc2 = factor(cnames[r$mc2], levels = cnames), w = r$w)
	scdb_add_cgraph(graph_id, tgCellGraph(gr, cnames))
}

# replaces the tgs_graph_cover_resample call inside mcell_coclust_from_graph_resamp (R/coclust_graph_cov.r:41);
# the wrapper around it (:13-58) is unchanged.  Seeds and index sets are drawn here, on the host.
tgs_graph_cover_resample = function(edges, knn, min_mc_size, cooling, burn_in, p_resamp, n_resamp, method = "full")
{
	if(!(method %in% c("full", "edges"))) stop("MC-ERR: method must be \"full\" or \"edges\"")
	nodes = levels(edges$col1)
	N = length(nodes)
	n_s = round(p_resamp * N)
	samples = sapply(1:n_resamp, function(i) sort(sample.int(N, n_s)))            # n_s x n_resamp, ascending
	seeds = floor(runif(n_resamp) * 2^52)
	# "edges": co-occurrence of graph edges only, O(edges) memory -- for cell counts whose N x N matrix cannot exist
	r = .Call(if(method == "full") "mcbr_coclust" else "mcbr_coclust_edges", N, as.integer(edges$col1), as.integer(edges$col2), as.numeric(edges$weight), as.integer(knn),
	          samples, seeds, as.integer(min_mc_size), as.numeric(cooling), as.integer(burn_in), PACKAGE = "mcb200")
	list(co_cluster = data.frame(node1 = factor(nodes[r$node1], levels = nodes),
	                             node2 = factor(nodes[r$node2], levels = nodes), cnt = r$cnt),
	     samples = r$samples)
}

# replaces tgstat's tgs_graph_cover at its two call sites, mcell_add_mc_from_graph (R/mc_graphcov.r:27) and
# mcell_mc_from_coclust_balanced (R/mc_from_coclust.r:247): data.frame(node, cluster), cluster 0 = outlier (R/mc_graphcov.r:28)
tgs_graph_cover = function(edges, min_mc_size, cooling, burn_in)
{
	nodes = levels(edges$col1)
	cl = .Call("mcbr_graph_cover", length(nodes), as.integer(edges$col1), as.integer(edges$col2), as.numeric(edges$weight),
	           as.integer(min_mc_size), as.numeric(cooling), as.integer(burn_in), floor(runif(1) * 2^52), PACKAGE = "mcb200")
	data.frame(node = nodes, cluster = cl, stringsAsFactors = F)
}

# several B200s in the box: options(mcb200.devices = 0:7) before the first call; mcell_add_cgraph_from_mat_bknn and
# tgs_graph_cover_resample then run on all of them (mcb_cgraph_bknn_multi / mcb_coclust_multi), nothing else changes
.onUnload = function(libpath) .Call("mcbr_shutdown", PACKAGE = "mcb200")
