"""Host-side mirror of the reference interface (metacell_b200/api.py): containers, scdb, parameters,
error behaviour -- everything that needs no GPU."""
import numpy as np
import pytest


def test_params_follow_reference_yaml():
    from metacell_b200 import api
    assert api.get_param("scm_k_nonz_exp") == 7                       # metacell_params.yaml:44
    assert api.get_param("scm_balance_graph_k_beta") == 3             # :50
    assert api.get_param("scm_balance_graph_k_expand") == 10          # :51
    assert api.get_param("scm_tgs_clust_cool") == 1.05 and api.get_param("scm_tgs_clust_burn_in") == 10
    assert api.get_param("mc_rseed") == 42
    api.set_param("mc_rseed", 7)
    assert api.get_param("mc_rseed") == 7
    api.set_param("mc_rseed", 42)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.get_param("nope")


def test_scdb_roundtrip_and_errors(tmp_path):
    from metacell_b200 import api
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.scdb_init(str(tmp_path / "missing"))
    api.scdb_init(str(tmp_path), force_reinit=True)
    cg = api.tgCellGraph({"mc1": np.array([0, 1]), "mc2": np.array([1, 2]), "w": np.array([1.0, 0.5])}, ["a", "b", "c"])
    api.scdb_add_cgraph("g1", cg)
    assert (tmp_path / "cgraph.g1.pkl").exists()                      # "<base>/<type>.<id>" naming, R/scdb.r:100-131
    api.scdb_init(str(tmp_path), force_reinit=True)                   # drops the cache -> reload from disk
    back = api.scdb_cgraph("g1")
    assert back.cell_names == ["a", "b", "c"] and np.array_equal(back.edges["mc2"], [1, 2])
    assert api.scdb_cgraph("none") is None
    coc = api.tgCoClust("g1", {"node1": np.array([0]), "node2": np.array([1]), "cnt": np.array([3])}, np.array([1, 1, 1]))
    api.scdb_add_coclust("c1", coc)
    assert api.scdb_coclust("c1").graph_id == "g1"
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgCoClust("g1", {"node1": [0], "node2": [1]}, [1])        # missing cnt column (R/coclust.r:39-41)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgCoClust("nope", {"node1": [0], "node2": [1], "cnt": [1]}, [1])


def test_cell_graph_container(capsys):
    from metacell_b200 import api
    cg = api.tgCellGraph({"mc1": np.array([0, 0]), "mc2": np.array([1, 2]), "w": np.array([1.0, 0.9])}, ["a", "b", "c", "d"])
    assert cg.nodes == ["a", "b", "c"]
    assert "missing 1 nodes" in capsys.readouterr().err              # R/cgraph.r:36-39
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.tgCellGraph({"mc1": [0], "mc2": [1]}, ["a", "b"])


def test_feature_rows_follow_gene_set_order():
    from metacell_b200 import api
    mat = api.tgScMat([0, 1, 2], [0, 3], [1.0, 2.0], genes=["g0", "g1", "g2", "g3"])
    gs = api.tgGeneSets({"g3": 1, "gX": 1, "g1": 2, "g0": 1})
    assert api._feature_rows(gs, mat).tolist() == [3, 1, 0]           # intersect(names(gset), rownames), R/gset.r:192-197
    with pytest.raises(api.McbError, match="MC-ERR"):
        api._feature_rows(api.tgGeneSets({"g0": 1}), mat)


def test_missing_objects_raise_reference_style_errors(tmp_path):
    from metacell_b200 import api
    api.scdb_init(str(tmp_path), force_reinit=True)
    with pytest.raises(api.McbError, match="MC-ERR: non existing gset"):
        api.mcell_add_cgraph_from_mat_bknn("m", "gs", "g", 100)
    with pytest.raises(api.McbError, match="MC-ERR: cell graph id"):
        api.mcell_coclust_from_graph_resamp("c", "g", 20, 0.75, 10)


def test_resample_index_sets():
    from metacell_b200 import api
    sets, seeds = api.resample_index_sets(1000, 0.75, 6, 42)
    assert sets.shape == (6, 750) and seeds.shape == (6,) and seeds.dtype == np.uint64
    assert np.all(np.diff(sets, axis=1) > 0) and sets.min() >= 0 and sets.max() < 1000
    s2, d2 = api.resample_index_sets(1000, 0.75, 6, 42)
    assert np.array_equal(sets, s2) and np.array_equal(seeds, d2)
    assert len({tuple(r) for r in sets}) == 6


def test_synth_shapes():
    from metacell_b200 import synth
    m = synth.make_umi_matrix(300, 200, n_types=5, median_umis=1000.0, seed=1)
    assert m.indptr.size == 301 and m.indices.max() < 200 and m.data.min() >= 1
    assert abs(np.median(m.col_sums()) / 1000.0 - 1.0) < 0.5


def test_coclust_filt_by_k_deg_matches_r_logic(tmp_path):
    """mcell_coclust_filt_by_k_deg against a literal loop restatement of R/coclust.r:58-84"""
    from metacell_b200 import api
    rng = np.random.default_rng(5)
    nn = 40
    pairs = [(a, b) for a in range(nn) for b in range(a + 1, nn) if rng.random() < 0.3]
    n1 = np.array([p[0] for p in pairs]); n2 = np.array([p[1] for p in pairs])
    cnt = rng.integers(1, 12, size=len(pairs))
    api.scdb_init(str(tmp_path), force_reinit=True)
    api.scdb_add_cgraph("g", api.tgCellGraph({"mc1": n1, "mc2": n2, "w": np.ones(len(pairs))}, [f"c{i}" for i in range(nn)]))
    api.scdb_add_coclust("coc", api.tgCoClust("g", {"node1": n1, "node2": n2, "cnt": cnt}, np.full(nn, 20)))
    K, alpha = 4, 2
    got = api.mcell_coclust_filt_by_k_deg("coc", K, alpha)
    levels = sorted(set(cnt.tolist()))
    thresh = {}
    for v in range(nn):
        mine = [c for a, b, c in zip(n1, n2, cnt) if a == v or b == v]
        if not mine:
            continue
        cum, t = 0, 0
        for lv in reversed(levels):                     # cumsum(rev(table row))
            cum += sum(1 for c in mine if c == lv)
            t += cum > K
        thresh[v] = t
    exp = np.array([(thresh[a] < c * alpha) or (thresh[b] < c * alpha) for a, b, c in zip(n1, n2, cnt)])
    assert np.array_equal(got, exp)
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mcell_coclust_filt_by_k_deg("nope", K, alpha)


def test_rollmedian_matches_oracle_and_gene_stat_oracle_on_tiny_case():
    """host logic of scm_gene_stat: the running-median helper against the loop restatement, and the dense oracle
    against hand-computed values on a 2 x 5 matrix (reference R/gstats.r:83-87, :113-120)"""
    from metacell_b200 import api
    from oracle import oracle_np
    rng = np.random.default_rng(0)
    x = rng.normal(size=400)
    assert np.array_equal(api._rollmedian(x, 101, -1.0, 2.0), oracle_np.rollmedian_fill(x, 101, -1.0, 2.0))
    raw = np.array([[0, 3, 1, 0, 6], [2, 2, 2, 2, 2]], dtype=np.float64)
    st = oracle_np.gene_stats_dense(raw, None, niche_quantile=0.4)       # round(5 * 0.4) = 2 cells
    assert np.array_equal(st["tot"], [10, 10])
    assert np.allclose(st["niche_stat"], [(10 + 9) / (10 + 10), (10 + 4) / (10 + 10)])
    assert np.allclose(st["var"], [np.var([0, 3, 1, 0, 6], ddof=1), 0.0])
    assert np.array_equal(st["is_on_count"], [3, 5])
    csum = raw.sum(axis=0)
    assert np.allclose(st["n_mean"], (raw * (1000.0 / csum)).mean(axis=1))
    assert np.isnan(st["sz_cor"][1]) and abs(st["sz_cor"][0] - np.corrcoef(raw[0], csum)[0, 1]) < 1e-15


def test_mc_confusion_mat_host_logic(tmp_path):
    """mcell_mc_confusion_mat is host tabulation (reference R/mc_confusion.r:10-45): edges between covered cells are
    counted per (metacell of source, metacell of target); cells outside the cover contribute nothing"""
    from metacell_b200 import api
    api.scdb_init(str(tmp_path), force_reinit=True)
    cells = [f"c{i}" for i in range(6)]
    edges = {"mc1": np.array([0, 0, 1, 2, 3, 4, 5]), "mc2": np.array([1, 2, 0, 3, 2, 5, 0]), "w": np.ones(7)}
    api.scdb_add_cgraph("g", api.tgCellGraph(edges, cells))
    api._scdb_add("mc", "m", {"mc": {"c0": 1, "c1": 1, "c2": 2, "c3": 2, "c4": 2}, "outliers": ["c5"]})
    confu = api.mcell_mc_confusion_mat("m", "g", 2)
    # c0->c1 (1,1) c0->c2 (1,2) c1->c0 (1,1) c2->c3 (2,2) c3->c2 (2,2); c4->c5 and c5->c0 touch the outlier
    assert np.array_equal(confu, np.array([[2, 1], [0, 2]]))
    api._scdb_add("mc", "bad", {"mc": {"zz": 1}, "outliers": []})
    with pytest.raises(api.McbError, match="MC-ERR"):
        api.mcell_mc_confusion_mat("bad", "g", 2)
    assert api.mcell_mc_confusion_mat("bad", "g", 2, ignore_mismatch=True).size == 0


def test_r_glue_type_checks_against_the_abi():
    """r/mcb200_glue.c cannot be built here (no R headers), but every mcb_* call in it is type-checked against
    include/mcb200.h by compiling it (-fsyntax-only) with declarations-only stubs of the R C API (tests/r_stub/)."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(root, "tests", "r_stub"),
                        "-I" + os.path.join(root, "include"), os.path.join(root, "r", "mcb200_glue.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every .Call target named in r/mcb200.R exists in the glue
    glue = open(os.path.join(root, "r", "mcb200_glue.c")).read()
    for name in set(re.findall(r'\.Call\("(mcbr_[a-z_]+)"', open(os.path.join(root, "r", "mcb200.R")).read())):
        assert re.search(r"\bSEXP " + name + r"\(", glue), name
