"""lock-step SearchKNN with the frontier, the batch-PIR bookkeeping and the clients' local caches on
the GPU (pm_search_*), against the CPU oracle lane by lane -- results, reach steps and the reference's success
accounting -- and against the host-driven path (PM_SEARCH_DEVICE=0), which must agree bit for bit."""
import os

import numpy as np
import pytest

from test_graphann_gpu import make_dataset

pytestmark = pytest.mark.gpu


@pytest.fixture()
def search_mode():
    saved = os.environ.get("PM_SEARCH_DEVICE")

    def set_(device):
        os.environ["PM_SEARCH_DEVICE"] = "1" if device else "0"

    yield set_
    if saved is None:
        os.environ.pop("PM_SEARCH_DEVICE", None)
    else:
        os.environ["PM_SEARCH_DEVICE"] = saved


def _oracle_lane(oracle, vec, graph, seed, start, queries, k, steps, par):
    from pacmann_b200.keys import mix64
    n, dim = vec.shape
    m = graph.shape[1]
    raw = oracle.pack_db(vec, graph)
    o_pir = oracle.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
    o_pir.preprocessing(key_seed=mix64(seed, 1), repl_seed=mix64(seed, 2), threads=4)
    return oracle.search_knn_private(o_pir, vec, graph, start, queries, k, steps, par)


@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("n,dim,m,lanes,nq,k,steps,par,integer", [
    (6000, 32, 8, 3, 11, 10, 10, 2, False),     # ragged last round (11 queries on 3 lanes)
    (6000, 32, 8, 2, 44, 10, 12, 2, False),     # every lane crosses its batch budget several times: Preprocessing() between steps
    (5000, 16, 16, 4, 16, 10, 8, 3, True),      # integer-valued vectors (SIFT-shaped): many equal distances -> heap / ranking tie rules
    (600, 8, 8, 2, 12, 5, 12, 2, False),        # tiny partitions: a sub-PIR budget can run out inside a round -> lanes leave the device path
    (20000, 64, 32, 2, 6, 20, 6, 3, False),     # degree 32, 16 partitions, 6 sub-queries per partition and call (the paper's geometry)
])
def test_device_lockstep_matches_oracle(oracle, search_mode, device, n, dim, m, lanes, nq, k, steps, par, integer):
    from pacmann_b200 import graphann
    search_mode(device)
    vec, graph = make_dataset(n, dim, m, 200 + n % 97, integer)
    queries = vec[np.random.default_rng(n + 1).integers(0, n, nq)] + np.float32(0.02 if not integer else 0.0)
    seeds = [900 + 13 * i for i in range(lanes)]
    group = graphann.make_client_group(vec, graph, lanes, seeds=seeds)
    before = graphann.DeviceSearchStats()
    ret, step = graphann.SearchKNNLockstep(group, queries, k, steps, par)
    after = graphann.DeviceSearchStats()
    if device:      # the device path really ran (no silent fall-back to the host-driven loop) ...
        assert after[1] + after[2] - before[1] - before[2] == nq and after[1] > before[1]
        if n == 600:
            assert after[2] > before[2]     # ... and in the tiny-partition case some lanes had to leave it for a while
    else:
        assert after == before
    for l in range(lanes):
        o_ret, o_step, stats = _oracle_lane(oracle, vec, graph, seeds[l], group[l].StartVertexIds(), queries[l::lanes], k, steps, par)
        assert (ret[l::lanes] == o_ret).all(), f"lane {l}: results differ from the oracle"
        assert (step[l::lanes] == o_step).all(), f"lane {l}: reach steps differ from the oracle"
        assert (group[l].totalQueryNum, group[l].succQueryNum) == (int(stats[0]), int(stats[1])), f"lane {l}: success accounting"
    # the group keeps working afterwards: a second call continues every client's state (caches, budgets)
    ret2, step2 = graphann.SearchKNNLockstep(group, queries[:lanes], k, steps, par)
    assert ret2.shape == (lanes, k)


def test_device_lockstep_second_call_continues_client_state(oracle, search_mode):
    """two SearchKNNLockstep calls == one oracle run over the concatenated queries (local caches and budgets carry over);
    repeated queries are served from the clients' caches"""
    from pacmann_b200 import graphann
    search_mode(True)
    n, dim, m, lanes, k, steps, par = 8000, 32, 8, 2, 10, 10, 2
    vec, graph = make_dataset(n, dim, m, 77)
    q1 = vec[np.random.default_rng(5).integers(0, n, 6)] + np.float32(0.02)
    queries = np.concatenate([q1, q1[:4], q1])          # repeats: the same vertices are fetched again -> cache hits
    seeds = [41, 42]
    group = graphann.make_client_group(vec, graph, lanes, seeds=seeds)
    a, sa = graphann.SearchKNNLockstep(group, queries[:6], k, steps, par)
    b, sb = graphann.SearchKNNLockstep(group, queries[6:], k, steps, par)
    ret, step = np.concatenate([a, b]), np.concatenate([sa, sb])
    for l in range(lanes):
        o_ret, o_step, stats = _oracle_lane(oracle, vec, graph, seeds[l], group[l].StartVertexIds(), queries[l::lanes], k, steps, par)
        assert (ret[l::lanes] == o_ret).all() and (step[l::lanes] == o_step).all(), f"lane {l}"
        assert (group[l].totalQueryNum, group[l].succQueryNum) == (int(stats[0]), int(stats[1]))
        assert group[l].PIR.serverQueries < group[l].totalQueryNum      # cache hits and drops never reach the server


def test_device_lockstep_benchmark_mode(search_mode):
    """-benchmark mode on the device path: random ids every step, nothing becomes known, all results -1"""
    from pacmann_b200 import graphann
    search_mode(True)
    n, dim, m, lanes = 3000, 32, 8, 2
    vec, graph = make_dataset(n, dim, m, 67)
    group = graphann.make_client_group(vec, graph, lanes, seeds=[3, 4], skipPrep=True)
    ret, step = graphann.SearchKNNLockstep(group, vec[:4], 5, 4, 2, benchmarking=True)
    assert (ret == -1).all() and (step == -1).all()
    for f in group:
        assert f.totalQueryNum == 2 * 4 * 2 * m
