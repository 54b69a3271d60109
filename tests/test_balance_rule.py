"""The only executable statement of the balancing rule in the reference tree is cormat_to_balanced_knn
(reference R/cgraph_knn_mcnorm.r:48-81); tgs_graph itself lives in tgstat, which is absent.  This file transliterates that
R function literally (base-R semantics of order / rank(ties = "average") / pmax / apply restated in numpy) and pins
oracle orc_balance to it wherever the two CAN agree, and records exactly where they differ:

  agree   the thresholded pair set: M - rank(i,j) * rank(j,i) > 0 with M = K^2 * 10  ==  s < K^2 * k_expand (strict),
          whenever both directions are listed (:66-68);
  agree   the first cut, "no more than 3K contribs per row, or incoming" (:70-71) == orc_balance's in-degree cut K * k_beta
          on every row whose cut is not inside a run of tied scores;
  agree   both degree caps (<= 3K per row after :71, <= K per column after :74);
  differ  absent ranks: R initialises every missing rank to K * k_alpha (:55-58), so a ONE-SIDED pair whose listed rank is
          below K passes the threshold in both directions; orc_balance (DECISION "absent => edge dropped", after the
          [RECALL] tgstat documentation: sym_rank is NA when either rank is missing) drops it;
  differ  ties: R's rank() gives tied scores their average rank, so a run of ties at a cut is kept or dropped as a whole;
          orc_balance breaks ties by the other node's index;
  differ  the second cut ranks the FIRST cut's residual ranks (:74 applies rank() to m_knn_io_rows), not the scores, so which
          K survive per column is not a function of the scores alone; orc_balance keeps the K best scores;
  differ  direction and weights: order(cormat[i, ]) is ascending (:61), i.e. the LOWEST values are ranked first, and the
          emitted weights are residual ranks 1..K rather than rank fractions.
"""
import numpy as np
import pytest

from oracle import oracle as O


def r_rank_average(x):
    """base-R rank(x) with ties.method = "average" (1-based)"""
    x = np.asarray(x, dtype=np.float64)
    order = np.argsort(x, kind="stable")
    ranks = np.empty(x.size, dtype=np.float64)
    sx = x[order]
    i = 0
    while i < x.size:
        j = i
        while j + 1 < x.size and sx[j + 1] == sx[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def cormat_to_balanced_knn(cormat, k_knn, k_alpha, k_beta):
    """Literal transliteration of reference R/cgraph_knn_mcnorm.r:48-81 (k_beta is unused there: 3 is hard-coded at :70).
    Returns (m_knn, m_knn_io after :68, m_knn_io_rows after :71, final m_knn_io after :74)."""
    N = cormat.shape[0]
    kp = min(k_knn * k_alpha, N)                                  # :53
    m_knn = np.full((N, N), float(kp))                            # :55-57
    for i in range(N):                                            # :59-62
        best_hits = np.argsort(cormat[i], kind="stable")[:kp]     # head(order(cormat[i,]), kp): ascending
        m_knn[i, best_hits] = np.arange(1, kp + 1)
    M = k_knn * k_knn * 10                                        # :65
    m_io = np.maximum(-m_knn * m_knn.T + M, 0)                    # :67
    A = N - k_knn * 3                                             # :69
    rows = np.stack([np.maximum(r_rank_average(m_io[i]) - A, 0) for i in range(N)])           # :70  t(apply(., 1, .))
    A = N - k_knn                                                 # :72
    final = np.stack([np.maximum(r_rank_average(rows[:, j]) - A, 0) for j in range(N)], axis=1)  # :73  apply(., 2, .)
    return m_knn, m_io, rows, final


def lists_like_r(cormat, kp):
    """the ranked lists the R function builds at :59-62, in the layout orc_balance takes (rank order, -1 padded)"""
    N = cormat.shape[0]
    col = np.full((N, kp), -1, np.int32)
    for i in range(N):
        col[i] = np.argsort(cormat[i], kind="stable")[:kp]
    return col


def _sym_matrix(N, seed):
    rng = np.random.default_rng(seed)
    a = rng.random((N, N))
    c = (a + a.T) / 2
    np.fill_diagonal(c, -1.0)            # ascending order(): the cell itself is ranked first, as tgs_cor_knn ranks self first
    return c


def _scores(col):
    """s(i,j) = rank(i,j) * rank(j,i) from ranked lists; 0 where either direction is missing"""
    N, kp = col.shape
    rank = np.zeros((N, N), np.int64)
    for i in range(N):
        rank[i, col[i]] = np.arange(1, kp + 1)
    return rank * rank.T, rank


@pytest.mark.parametrize("N,K,seed", [(30, 3, 1), (40, 4, 2), (20, 2, 5)])
def test_threshold_set_and_first_cut_agree_when_every_pair_is_listed(N, K, seed):
    """k_alpha = 10 and N <= 10 K: every cell lists every cell, no rank is absent"""
    c = _sym_matrix(N, seed)
    m_knn, m_io, rows, final = cormat_to_balanced_knn(c, K, 10, 3)
    col = lists_like_r(c, min(10 * K, N))
    s, rank = _scores(col)
    assert np.array_equal(rank, m_knn.astype(np.int64))
    # (1) threshold: strict, exactly orc_balance's default
    assert np.array_equal(m_io > 0, s < K * K * 10)
    # (2) first cut against orc_balance with the out-degree cut out of the way: K_out = N cannot bind, in cut = 3 K
    #     orc_balance(K' = N ...) would change the threshold, so the comparison goes through the final graph instead:
    deg, dst, sc = O.balance(col, K, 10, 3, keep_self=True)
    kept_r1 = rows > 0                                   # [target row i][source j] after "no more than 3K per row"
    thr_ok = s < K * K * 10
    for i in range(N):
        cand = np.flatnonzero(thr_ok[i])
        keys = sorted((int(s[i, j]), int(j)) for j in cand)
        if len(keys) > 3 * K and keys[3 * K - 1][0] == keys[3 * K][0]:
            continue                                    # the cut falls inside a run of ties: R keeps / drops the run as a whole
        want = {j for _, j in keys[:3 * K]}
        assert set(np.flatnonzero(kept_r1[i]).tolist()) == want, i
    # every edge orc_balance emits (source a -> target b) survived the first cut of ITS target row in the R statement,
    # unless that row's cut is tied
    for a in range(N):
        for b in dst[a, :deg[a]]:
            keys = sorted((int(s[b, j]), int(j)) for j in np.flatnonzero(thr_ok[b]))
            tied = len(keys) > 3 * K and keys[3 * K - 1][0] == keys[3 * K][0]
            assert tied or kept_r1[b, a], (a, b)
    # (3) degree caps of both statements
    for i in range(N):
        vals = np.sort(m_io[i])[::-1]
        no_tie = 3 * K >= N or vals[3 * K - 1] != vals[3 * K]
        if no_tie:
            assert kept_r1[i].sum() <= 3 * K
    assert (deg <= K).all()
    indeg = np.bincount(np.concatenate([dst[a, :deg[a]] for a in range(N)]), minlength=N)
    assert (indeg <= 3 * K).all()
    for j in range(N):
        v = np.sort(rows[:, j])[::-1]
        if K >= N or v[K - 1] != v[K]:
            assert (final[:, j] > 0).sum() <= K


def test_absent_ranks_are_where_the_two_statements_differ():
    """N > K * k_alpha: lists are partial.  R gives a missing rank the value K * k_alpha, so a one-sided pair with listed rank
    r passes iff r * K * k_alpha < K^2 * 10, i.e. r < K (k_alpha = 10); orc_balance drops every one-sided pair.  On mutual
    pairs the two thresholded sets are identical."""
    N, K = 60, 3
    c = _sym_matrix(N, 9)
    m_knn, m_io, rows, final = cormat_to_balanced_knn(c, K, 10, 3)
    kp = 10 * K
    col = lists_like_r(c, kp)
    s, rank = _scores(col)
    listed = rank > 0
    mutual = listed & listed.T
    assert np.array_equal((m_io > 0) & mutual, (s < K * K * 10) & mutual)
    one_sided = listed ^ listed.T
    r_listed = np.where(listed, rank, rank.T)
    assert np.array_equal((m_io > 0) & one_sided, one_sided & (r_listed < K))
    assert ((m_io > 0) & ~listed & ~listed.T).sum() == 0                  # kp * kp >= M: never both absent
    # orc_balance never emits a one-sided pair
    deg, dst, _ = O.balance(col, K, 10, 3, keep_self=True)
    for a in range(N):
        for b in dst[a, :deg[a]]:
            assert mutual[a, b]


def test_second_cut_ranks_residual_ranks_not_scores():
    """:74 applies rank() to the first cut's residual ranks.  Two entries of one column with scores s1 < s2 can therefore come
    out in the other order when they sit in rows of different crowding -- the reason the R function is read as a statement
    of the rule (threshold, 3K in, K out), not of which K survive."""
    found = False
    for seed in range(40):
        N, K = 30, 3
        c = _sym_matrix(N, 100 + seed)
        m_knn, m_io, rows, final = cormat_to_balanced_knn(c, K, 10, 3)
        for j in range(N):
            alive = np.flatnonzero(rows[:, j] > 0)
            for a in alive:
                for b in alive:
                    if m_io[a, j] > m_io[b, j] and rows[a, j] < rows[b, j]:      # better score, worse residual rank
                        found = True
        if found:
            break
    assert found
