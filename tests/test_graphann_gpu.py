"""graphann over the B200 path: L2Dist (TestDistance, graphann_test.go:15-58), SearchKNN against the oracle in
non-private and private mode, and the private-search.go wire format."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_dataset(n, dim, m, seed, integer=False):
    rng = np.random.default_rng(seed)
    if integer:   # SIFT-shaped: integers 0..255 stored as f32 (graphann/loader.go:47-51)
        vec = rng.integers(0, 256, (n, dim)).astype(np.float32)
    else:
        vec = (rng.standard_normal((n, dim)) * np.linspace(0.8, 0.3, dim)).astype(np.float32)
    # a searchable graph: half near neighbours in a 1-d projection, half random (no self loops, duplicates allowed)
    order = np.argsort(vec[:, 0])
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    graph = np.zeros((n, m), np.int32)
    for j in range(m // 2):
        off = (j // 2 + 1) * (1 if j % 2 == 0 else -1)
        graph[:, j] = order[np.clip(pos + off, 0, n - 1)]
    graph[:, m // 2:] = rng.integers(0, n, (n, m - m // 2))
    self_loop = graph == np.arange(n)[:, None]
    graph[self_loop] = (graph[self_loop] + 1) % n
    return vec, graph


def test_distance(oracle):
    """TestDistance: |L2Dist - L2DistSIMD| < 1e-4 at dim 128; here L2Dist is bit-identical to the oracle's."""
    from pacmann_b200 import graphann
    rng = np.random.default_rng(61)
    for _ in range(50):
        v1, v2 = rng.random(128, dtype=np.float32), rng.random(128, dtype=np.float32)
        truth = oracle.l2dist(v1, v2)
        got = graphann.L2Dist(v1, v2)
        assert abs(float(truth) - float(got)) < 1e-4          # the reference's tolerance
        assert np.float32(got).view(np.uint32) == np.float32(truth).view(np.uint32)   # ours: 0 ulp


@pytest.mark.parametrize("integer", [False, True])
def test_search_knn_nonprivate_matches_oracle(oracle, integer):
    from pacmann_b200 import graphann
    n, dim, m = 4000, 64, 16
    vec, graph = make_dataset(n, dim, m, 62, integer)
    queries = vec[np.random.default_rng(63).integers(0, n, 20)] + np.float32(0.01)
    f = graphann.GraphANNFrontend(vec, graph, private=False)
    f.Preprocess()
    start = f.StartVertexIds()
    assert (start == np.arange(int(np.sqrt(n)))).all()          # BasicGraphInfo.GetStartVertex (search.go:51-65)
    ret, step = f.SearchKNNBatch(queries, 10, 12, 3)
    o_ret, o_step = oracle.search_knn_basic(vec, graph, start, queries, 10, 12, 3)
    assert (ret == o_ret).all() and (step == o_step).all()


def test_private_search_matches_oracle_and_nonprivate(oracle):
    from pacmann_b200 import graphann
    from pacmann_b200.keys import mix64
    n, dim, m = 6000, 32, 8
    vec, graph = make_dataset(n, dim, m, 64)
    queries = vec[np.random.default_rng(65).integers(0, n, 12)] + np.float32(0.02)
    seed = 66
    f = graphann.GraphANNFrontend(vec, graph, private=True, seed=seed)
    f.Preprocess()
    start = f.StartVertexIds()
    assert len(set(start.tolist())) == int(np.sqrt(n))
    # SearchKNNBatch is a plain loop (search.go:236-245): run it query by query to see which queries had a failed fetch
    rets, steps, clean = [], [], []
    for q in queries:
        t0, s0 = f.totalQueryNum, f.succQueryNum
        r, s = f.SearchKNN(q, 10, 10, 2)
        rets.append(r)
        steps.append(s)
        clean.append(f.totalQueryNum - t0 == f.succQueryNum - s0)     # every fetched entry was the true row
    ret, step = np.stack(rets), np.stack(steps)

    # oracle: same DB packing, same seeds, same start vertices
    raw = oracle.pack_db(vec, graph)
    o_pir = oracle.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
    o_pir.preprocessing(key_seed=mix64(seed, 1), repl_seed=mix64(seed, 2), threads=4)
    o_ret, o_step, stats = oracle.search_knn_private(o_pir, vec, graph, start, queries, 10, 10, 2)
    assert (ret == o_ret).all() and (step == o_step).all()
    assert (f.totalQueryNum, f.succQueryNum) == (int(stats[0]), int(stats[1]))
    assert f.succQueryNum > 0.5 * f.totalQueryNum

    # non-private traversal over the same start vertices
    g = graphann.GraphANNFrontend(vec, graph, private=True, nonPrivateMode=True, skipPrep=True, seed=seed)
    g.Preprocess()
    assert (g.StartVertexIds() == start).all()
    np_ret, _ = g.SearchKNNBatch(queries, 10, 10, 2)
    b_ret, _ = oracle.search_knn_basic(vec, graph, start, queries, 10, 10, 2)
    assert (np_ret == b_ret).all()
    # a private search whose every fetch returned the true row walks exactly the non-private path (batch drops and
    # failed sub-queries are the only source of divergence): identical result for each such query
    for i, ok in enumerate(clean):
        if ok:
            assert (ret[i] == np_ret[i]).all(), f"query {i}: clean private search differs from the non-private one"
    if f.succQueryNum == f.totalQueryNum:
        assert (ret == np_ret).all()


def test_benchmark_mode_issues_random_queries(oracle):
    """-benchmark mode (private-search.go:85,191; search.go:155-159,182-185): DummyPreprocessing, random ids,
    answers discarded; result is all -1."""
    from pacmann_b200 import graphann
    n, dim, m = 3000, 32, 8
    vec, graph = make_dataset(n, dim, m, 67)
    f = graphann.GraphANNFrontend(vec, graph, private=True, skipPrep=True, seed=3)
    f.Preprocess()
    ret, step = f.SearchKNNBatch(vec[:3], 5, 4, 2, benchmarking=True)
    assert (ret == -1).all() and (step == -1).all()
    assert f.totalQueryNum == 3 * 4 * 2 * m
    # every slot is a server query unless its index repeated (local cache) or no hint matched
    assert 0.9 * 192 <= f.PIR.serverQueries <= 3 * 4 * 2 * m
    assert f.PIR.serverLaunches == 3 * 4


@pytest.mark.parametrize("lanes,nq,steps", [(3, 11, 10), (2, 44, 12)])
def test_lockstep_client_group_matches_oracle(oracle, lanes, nq, steps):
    """L independent clients whose hint tables live in ONE pm_client, searched in lock step (one device
    call per step for all lanes).  Query i goes to lane i % L and every lane must return exactly what the CPU oracle
    returns for a client with that lane's seed running its queries alone -- also across the batch budget
    (second case: each lane re-preprocesses several times, and the call before that runs outside the group)."""
    from pacmann_b200 import graphann
    from pacmann_b200.keys import mix64
    n, dim, m = 6000, 32, 8
    vec, graph = make_dataset(n, dim, m, 164)
    queries = vec[np.random.default_rng(165).integers(0, n, nq)] + np.float32(0.02)
    seeds = [300 + 7 * i for i in range(lanes)]
    group = graphann.make_client_group(vec, graph, lanes, seeds=seeds)
    ret, step = graphann.SearchKNNLockstep(group, queries, 10, steps, 2)
    raw = oracle.pack_db(vec, graph)
    for l in range(lanes):
        o_pir = oracle.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
        o_pir.preprocessing(key_seed=mix64(seeds[l], 1), repl_seed=mix64(seeds[l], 2), threads=4)
        o_ret, o_step, stats = oracle.search_knn_private(o_pir, vec, graph, group[l].StartVertexIds(), queries[l::lanes], 10, steps, 2)
        assert (ret[l::lanes] == o_ret).all() and (step[l::lanes] == o_step).all(), f"lane {l}"
        assert (group[l].totalQueryNum, group[l].succQueryNum) == (int(stats[0]), int(stats[1]))
    # a lane keeps working on its own afterwards (same pm_client, its own parts)
    solo_ret, _ = group[lanes - 1].SearchKNNBatch(queries[:2], 10, steps, 2)
    assert solo_ret.shape == (2, 10)


def test_l2_idpairs_matches_oracle(oracle):
    """pm_l2_idpairs: distances between rows of the resident table addressed by id pairs, bit-identical to L2Dist."""
    from pacmann_b200 import cabi
    rng = np.random.default_rng(171)
    n, dim = 500, 40
    vec = (rng.standard_normal((n, dim)) * 3).astype(np.float32)
    db = cabi.DB(vec.view(np.uint64).reshape(n, dim // 2))
    a, b = rng.integers(0, n, 777), rng.integers(0, n, 777)
    a[5], b[9] = -1, n                                        # out of range -> +inf
    got = cabi.l2_idpairs(db, dim, a, b)
    for p in range(777):
        if a[p] < 0 or b[p] >= n:
            assert np.isinf(got[p])
        else:
            assert got[p].view(np.uint32) == oracle.l2dist(vec[a[p]], vec[b[p]]).view(np.uint32)
    db.close()


@pytest.mark.parametrize("integer", [False, True])
def test_robust_prune_batch_matches_oracle(oracle, integer):
    """robustPrune (build_graph.go:169-236) with all its L2Dist calls served by one launch.  Same
    neighbour lists as the oracle restatement, vertex by vertex -- on integer (SIFT-shaped) data with many equal
    distances too, where the stated tie rule (candidate order) matters."""
    from pacmann_b200 import graphann
    rng = np.random.default_rng(172)
    n, dim, m, k = 3000, 32, 16, 40
    vec, _ = make_dataset(n, dim, 8, 173, integer=integer)
    if integer:
        vec = np.floor(vec / 64).astype(np.float32)           # few distinct values: ties everywhere
    us = rng.integers(0, n, 200)
    cand = np.stack([rng.choice(n, k, replace=False) for _ in us])
    cand[3, 7] = cand[3, 2]                                   # a duplicated candidate
    for alpha in (1.0, 1.2):
        got = graphann.RobustPruneBatch(vec, us, cand, m, alpha)
        for i, u in enumerate(us):
            want = oracle.robust_prune(vec, u, cand[i], m, alpha)
            assert len(got[i]) == len(want) and (got[i] == want).all(), (i, alpha)
    few = graphann.RobustPruneBatch(vec, us[:5], cand[:5, :m], m, 1.2)      # len(candidates) <= m: returned unchanged
    assert all((few[i] == cand[i, :m]).all() for i in range(5))


def test_evaluate_graph_quality(oracle):
    """EvaluateGraphQuality (build_graph.go:776-817): random dataset vertices searched in their own graph."""
    from pacmann_b200 import graphann
    n, dim, m = 4000, 32, 16
    vec, graph = make_dataset(n, dim, m, 174)
    hit_rate, avg_steps = graphann.EvaluateGraphQuality(vec, graph, numQueries=50, seed=1)
    targets = np.random.default_rng(1).integers(0, n, 50)
    f_start = graphann.GraphANNFrontend(vec, graph)
    f_start.Preprocess()
    o_ret, o_step = oracle.search_knn_basic(vec, graph, f_start.StartVertexIds(), vec[targets], 20, 20, 2)
    o_hit = o_ret[:, 0] == targets
    assert abs(hit_rate - o_hit.mean()) < 1e-12
    if o_hit.any():
        assert abs(avg_steps - o_step[o_hit, 0].mean()) < 1e-9


def test_lockstep_over_plain_frontends_and_benchmark_mode(oracle):
    """SearchKNNLockstep also accepts lanes that cannot share a device call (BasicGraphInfo: every lane fetches for
    itself) and the -benchmark mode of a client group (random ids, answers discarded): same results as lane by lane."""
    from pacmann_b200 import graphann
    n, dim, m = 3000, 32, 8
    vec, graph = make_dataset(n, dim, m, 181)
    queries = vec[np.random.default_rng(182).integers(0, n, 7)] + np.float32(0.02)
    lanes = [graphann.GraphANNFrontend(vec, graph) for _ in range(2)]
    for f in lanes:
        f.Preprocess()
    ret, step = graphann.SearchKNNLockstep(lanes, queries, 10, 8, 2)
    for l in range(2):
        ref = graphann.GraphANNFrontend(vec, graph)
        ref.Preprocess()
        want, wstep = ref.SearchKNNBatch(queries[l::2], 10, 8, 2)
        assert (ret[l::2] == want).all() and (step[l::2] == wstep).all()
    group = graphann.make_client_group(vec, graph, 3, seeds=[5, 6, 7], skipPrep=True)
    bret, bstep = graphann.SearchKNNLockstep(group, queries[:6], 5, 4, 2, benchmarking=True)
    assert (bret == -1).all() and (bstep == -1).all()
    assert [g.totalQueryNum for g in group] == [2 * 4 * 2 * m] * 3
