"""GPU-resident client (pm_client_*): hint tables in HBM, hint search / refresh on the GPU.
Responses and the complete client state must equal the sequential CPU oracle's."""
import numpy as np
import pytest

from test_pianopir_gpu import assert_same_state
from util import splitmix_db

pytestmark = pytest.mark.gpu


def make(DBSize, E, BatchSize, fail, seed, key_seed, repl_seed, oracle, resident=True):
    from pacmann_b200 import pianopir
    rawDB = splitmix_db(DBSize, E, seed=seed)
    PIR = pianopir.NewSimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB, fail)
    PIR.SetSeeds(key_seed, repl_seed)
    if resident:
        PIR.EnableResidentClient()
    o_pir = oracle.SimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB.reshape(-1), fail)
    return rawDB, PIR, o_pir


@pytest.mark.parametrize("E", [112, 16, 6, 7])
def test_resident_preprocessing_tables_equal_oracle(oracle, E):
    rawDB, PIR, o_pir = make(30001, E, 32, 8, 71, 72, 73, oracle)
    PIR.Preprocessing()
    o_pir.preprocessing(72, 73, threads=8)
    for i in (0, 5, 15):
        assert_same_state(PIR.subPIR(i), o_pir.sub(i))


def test_resident_queries_equal_oracle_and_reference_assertions(oracle):
    rawDB, PIR, o_pir = make(50000, 112, 32, 8, 74, 75, 76, oracle)
    PIR.Preprocessing()
    o_pir.preprocessing(75, 76, threads=8)
    rng = np.random.default_rng(77)
    n_ok = 0
    for it in range(40):
        n = 96 if it % 2 else 32
        batch = rng.integers(0, 50000, n).astype(np.uint64)
        if it % 4 == 0:
            batch[3] = batch[1]                       # duplicate: second one is a cache hit
        if it % 7 == 0:
            batch[:20] = rng.integers(0, 3125, 20)    # crowd one partition: surplus dropped -> zeros
        resp, _ = PIR.Query(batch)
        want = o_pir.query(batch)
        assert (resp == want).all(), f"batch {it}"
        for j in range(n):                            # TestBatchPIRPerf's check (pir_test.go:256-261)
            assert (resp[j] == 0).all() or (resp[j] == rawDB[batch[j]]).all()
            n_ok += int((resp[j] == rawDB[batch[j]]).all())
    assert n_ok > 1000
    for i in (0, 7, 15):
        assert_same_state(PIR.subPIR(i), o_pir.sub(i))
        assert PIR.subPIR(i).client("FinishedQueryNum") == o_pir.sub(i).get("finished_query_num")
    assert PIR.serverLaunches == 40 and PIR.FinishedBatchNum == o_pir.finished_batch_num


def test_resident_batch_pir_basic(oracle):
    """TestBatchPIRBasic (pir_test.go:60-202) on the resident client."""
    from pacmann_b200 import pianopir
    DBSize, E, BatchSize = 1000000, 16, 32
    rawDB = np.repeat(np.arange(DBSize, dtype=np.uint64)[:, None], E, axis=1)
    PIR = pianopir.NewSimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB, 20)
    PIR.SetSeeds(81, 82)
    PIR.EnableResidentClient()
    PIR.Preprocessing()
    cfg = PIR.Config()
    rng = np.random.default_rng(83)
    one = np.array([i * cfg.PartitionSize + int(rng.integers(0, cfg.PartitionSize)) for i in range(cfg.PartitionNum)], np.uint64)
    r, _ = PIR.Query(one)
    assert (r == rawDB[one]).all()
    four = np.array([i * cfg.PartitionSize + int(rng.integers(0, cfg.PartitionSize)) for i in range(cfg.PartitionNum) for _ in range(4)], np.uint64)
    r, _ = PIR.Query(four)
    assert (r == rawDB[four]).all()
    crowd = rng.choice(cfg.PartitionSize, BatchSize, replace=False).astype(np.uint64)
    r, _ = PIR.Query(crowd)
    assert (r[:2] == rawDB[crowd[:2]]).all() and (r[2:] == 0).all()


def test_resident_budget_and_redo_preprocessing(oracle):
    rawDB, PIR, o_pir = make(1600, 4, 8, 8, 84, 85, 86, oracle)     # 4 partitions of 400 rows, MaxQueryNum 119
    PIR.Preprocessing()
    o_pir.preprocessing(85, 86)
    maxq = PIR.subPIR(0).client("MaxQueryNum")
    rng = np.random.default_rng(87)
    redone = False
    for it in range(maxq + 10):
        batch = rng.integers(0, 1600, 8).astype(np.uint64)
        before = PIR.QueriesMadeInPartition
        resp, _ = PIR.Query(batch)
        assert (resp == o_pir.query(batch)).all(), f"batch {it}"
        redone = redone or PIR.QueriesMadeInPartition < before
    assert redone
    for i in range(4):
        assert_same_state(PIR.subPIR(i), o_pir.sub(i))


def test_resident_sub_pir_budget_exhaustion_mid_batch(oracle):
    """pir.go:527-530 inside a batch call: a crowd of real queries drives one sub-PIR to MaxQueryNum."""
    rawDB, PIR, o_pir = make(1600, 4, 8, 40, 88, 89, 90, oracle)
    PIR.Preprocessing()
    o_pir.preprocessing(89, 90)
    rng = np.random.default_rng(91)
    for it in range(12):
        # 64 indices, all in partition 0 -> 16 real queries to sub-PIR 0 per call (budget 119, batch budget 117)
        batch = rng.choice(400, 64, replace=False).astype(np.uint64)
        resp, _ = PIR.Query(batch)
        assert (resp == o_pir.query(batch)).all(), f"batch {it}"
    assert_same_state(PIR.subPIR(0), o_pir.sub(0))


def test_resident_dummy_preprocessing(oracle):
    rawDB, PIR, o_pir = make(8000, 16, 8, 8, 92, 93, 94, oracle)
    PIR.DummyPreprocessing()
    o_pir.dummy_preprocessing(93)
    batch = np.random.default_rng(95).integers(0, 8000, 16).astype(np.uint64)
    resp, _ = PIR.Query(batch)
    assert (resp == o_pir.query(batch)).all()
    assert_same_state(PIR.subPIR(1), o_pir.sub(1))


@pytest.mark.parametrize("resident", [True, False])
def test_private_search_both_client_modes(oracle, resident):
    from pacmann_b200 import graphann
    from pacmann_b200.keys import mix64
    from test_graphann_gpu import make_dataset
    n, dim, m = 6000, 32, 8
    vec, graph = make_dataset(n, dim, m, 96)
    queries = vec[np.random.default_rng(97).integers(0, n, 10)] + np.float32(0.02)
    f = graphann.GraphANNFrontend(vec, graph, private=True, seed=98, resident=resident)
    f.Preprocess()
    start = f.StartVertexIds()
    ret, step = f.SearchKNNBatch(queries, 10, 10, 2)
    raw = oracle.pack_db(vec, graph)
    o_pir = oracle.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
    o_pir.preprocessing(key_seed=mix64(98, 1), repl_seed=mix64(98, 2), threads=4)
    o_ret, o_step, stats = oracle.search_knn_private(o_pir, vec, graph, start, queries, 10, 10, 2)
    assert (ret == o_ret).all() and (step == o_step).all()
    assert (f.totalQueryNum, f.succQueryNum) == (int(stats[0]), int(stats[1]))
