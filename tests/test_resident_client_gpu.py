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


def test_client_lanes_at_the_c_abi(oracle):
    """pm_client_* with L x PartitionNum parts, straight through ctypes: two lanes in ONE pm_client answering one mixed
    call (pm_client_query_batch_l2m, per-query vector ids, page-locked result buffer from pm_host_alloc) must give what two
    separate single-lane clients give (pm_client_query_batch_l2), entry for entry, status for status, distance for
    distance -- and leave identical hint tables behind."""
    import ctypes as C
    from pacmann_b200 import cabi
    from util import splitmix_db
    L = cabi.lib()
    n, E, dim = 4000, 20, 24
    rows = splitmix_db(n, E, seed=5)
    fl = np.random.default_rng(6).standard_normal((n, dim)).astype(np.float32)
    rows[:, :dim // 2] = fl.view(np.uint64)                       # entries start with a dim-float vector
    db = cabi.DB(rows)
    o = oracle.PianoPIR(n, E * 8, rows.reshape(-1), 8)
    geo = np.array([0, n, o.chunk_size, o.set_size, o.primary_hint_num, o.max_query_per_chunk, o.max_query_num], np.uint64)
    parts2 = np.concatenate([geo, geo])
    keys = [oracle.derive_key(11, 0, 2, i) for i in range(2)]
    rk = np.concatenate([cabi.expand_key(k) for k in keys]).astype(np.uint32)
    seeds = np.array([101, 202], np.uint64)

    def create(parts, nparts):
        h = C.c_void_p()
        cabi.check(L.pm_client_create(db.h, parts.ctypes.data_as(C.c_void_p), nparts, C.byref(h)))
        return h

    both = create(parts2, 2)
    ids = np.array([0, 1], np.uint32)
    cabi.check(L.pm_client_preprocess(both, ids.ctypes.data_as(C.c_void_p), 2, rk.ctypes.data_as(C.c_void_p), seeds.ctypes.data_as(C.c_void_p), 0))
    solo = []
    for i in range(2):
        h = create(geo, 1)
        z = np.array([0], np.uint32)
        cabi.check(L.pm_client_preprocess(h, z.ctypes.data_as(C.c_void_p), 1, rk[44 * i:44 * (i + 1)].ctypes.data_as(C.c_void_p),
                                          seeds[i:i + 1].ctypes.data_as(C.c_void_p), 0))
        solo.append(h)

    qdt = np.dtype([("part", np.uint32), ("kind", np.uint32), ("idx", np.uint64), ("dseed", np.uint64), ("dctr", np.uint64)])
    rng = np.random.default_rng(7)
    qv = rng.standard_normal((2, dim)).astype(np.float32)
    pbuf = C.c_void_p()
    nq = 24
    cabi.check(L.pm_host_alloc(C.byref(pbuf), nq * E * 8))
    out_m = np.ctypeslib.as_array(C.cast(pbuf, C.POINTER(C.c_uint64)), shape=(nq, E))
    for rnd in range(6):
        q = np.zeros(nq, qdt)
        q["part"] = rng.integers(0, 2, nq)
        q["kind"] = (rng.random(nq) < 0.85).astype(np.uint32)
        q["idx"] = rng.permutation(n)[:nq]                          # no index twice in a call
        q["dseed"], q["dctr"] = 99, np.arange(nq) * 1000 + rnd * 100000
        vid = q["part"].astype(np.uint32)
        st_m, di_m = np.zeros(nq, np.int32), np.zeros(nq, np.float32)
        cabi.check(L.pm_client_query_batch_l2m(both, q.ctypes.data_as(C.c_void_p), nq, pbuf, st_m.ctypes.data_as(C.c_void_p),
                                               qv.ctypes.data_as(C.c_void_p), 2, vid.ctypes.data_as(C.c_void_p), dim,
                                               di_m.ctypes.data_as(C.c_void_p)))
        for i in range(2):
            sel = np.nonzero(q["part"] == i)[0]
            qs = q[sel].copy()
            qs["part"] = 0
            o_s, st_s, di_s = np.zeros((len(sel), E), np.uint64), np.zeros(len(sel), np.int32), np.zeros(len(sel), np.float32)
            cabi.check(L.pm_client_query_batch_l2(solo[i], qs.ctypes.data_as(C.c_void_p), len(sel), o_s.ctypes.data_as(C.c_void_p),
                                                  st_s.ctypes.data_as(C.c_void_p), qv[i].ctypes.data_as(C.c_void_p), dim,
                                                  di_s.ctypes.data_as(C.c_void_p)))
            assert (out_m[sel] == o_s).all() and (st_m[sel] == st_s).all()
            assert (di_m[sel].view(np.uint32) == di_s.view(np.uint32)).all()
            real_ok = (qs["kind"] == 1) & (st_s == 0)
            assert (o_s[real_ok] == rows[qs["idx"][real_ok]]).all()
    P, B = int(o.primary_hint_num), int(o.set_size * o.max_query_per_chunk)
    for table, words in ((0, P), (1, P * E), (2, P), (5, B), (6, B * E), (7, int(o.set_size)), (8, 1)):
        for i in range(2):
            a, b = np.zeros(words, np.uint64), np.zeros(words, np.uint64)
            cabi.check(L.pm_client_download(both, i, table, a.ctypes.data_as(C.c_void_p), words))
            cabi.check(L.pm_client_download(solo[i], 0, table, b.ctypes.data_as(C.c_void_p), words))
            assert (a == b).all(), (table, i)
    cabi.check(L.pm_host_free(pbuf))
    for h in [both] + solo:
        L.pm_client_destroy(h)
    db.close()
