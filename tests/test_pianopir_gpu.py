"""The reference's own PIR tests (pianopir/pir_test.go) re-expressed over the B200 path, plus state parity of the
whole client against the CPU oracle (same injected keys and counter-based draws on both sides)."""
import numpy as np
import pytest

from util import splitmix_db

pytestmark = pytest.mark.gpu

TABLES = ["primaryShortTag", "primaryParity", "primaryProgramPoint", "replacementIdx", "replacementVal",
          "backupShortTag", "backupParity", "QueryHistogram"]
ORACLE_NAME = dict(primaryShortTag="primary_short_tag", primaryParity="primary_parity", primaryProgramPoint="primary_program_point",
                   replacementIdx="replacement_idx", replacementVal="replacement_val", backupShortTag="backup_short_tag",
                   backupParity="backup_parity", QueryHistogram="query_histogram")


def assert_same_state(p, o_pir):
    for t in TABLES:
        a, b = p.table(t), o_pir.table(ORACLE_NAME[t])
        assert a.shape == b.shape and (a == b).all(), t


def test_pir_basic(oracle):
    """TestPIRBasic (pir_test.go:9-58): every one of MaxQueryNum random real queries returns rawDB[idx]."""
    from pacmann_b200 import pianopir
    from pacmann_b200.keys import derive_key, mix64
    DBSize, DBEntrySize = 18750, 4
    rawDB = splitmix_db(DBSize, DBEntrySize, seed=21)
    PIR = pianopir.NewPianoPIR(DBSize, DBEntrySize * 8, rawDB, 40)
    cfg = PIR.Config()
    assert (cfg.ChunkSize, cfg.SetSize, PIR.client("primaryHintNum"), PIR.client("MaxQueryNum")) == (512, 40, 14848, 1347)
    PIR.SetSeeds(key_seed=5, epoch=0, repl_seed=6)
    PIR.Preprocessing()

    o_pir = oracle.PianoPIR(DBSize, DBEntrySize * 8, rawDB.reshape(-1), 40)
    o_pir.preprocessing(derive_key(5, 0, 1, 0), repl_seed=mix64(6, 0))
    assert (PIR.long_key() == o_pir.long_key()).all()
    assert_same_state(PIR, o_pir)

    rng = np.random.default_rng(22)
    for i in range(PIR.client("MaxQueryNum")):
        idx = int(rng.integers(0, DBSize))
        query, err = PIR.Query(idx, True)
        assert err == 0, f"PIR.Query({idx}) failed: {err}"
        assert (query == rawDB[idx]).all()
        o_q, o_rc = o_pir.client_query(idx, True)
        assert o_rc == 0 and (o_q == query).all()
    assert_same_state(PIR, o_pir)     # refreshed hints, program points, histogram: identical after 1347 queries
    assert PIR.client("FinishedQueryNum") == o_pir.get("finished_query_num")


def test_pir_dummy_out_of_range_and_server_paths(oracle):
    from pacmann_b200 import pianopir
    DBSize, E = 3000, 8
    rawDB = splitmix_db(DBSize, E, seed=23)
    PIR = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 8)
    PIR.SetSeeds(1, 0, 2)
    PIR.Preprocessing()
    q, err = PIR.Query(0, False)            # dummy query: zero entry, nil error (pir.go:363-371)
    assert err == 0 and (q == 0).all()
    q, err = PIR.Query(DBSize + 5, True)    # out of range (the Go code log.Fatalf's; here an error code)
    assert err == pianopir.ERR_OUT_OF_RANGE and (q == 0).all()
    # NonePrivateQuery: in range, in the padding, past the padding (pir.go:41-62)
    cfg = PIR.Config()
    v, e = PIR.NonePrivateQuery(17)
    assert e == 0 and (v == rawDB[17]).all()
    v, e = PIR.NonePrivateQuery(DBSize)
    assert (v == 0).all() and e == (0 if DBSize < cfg.ChunkSize * cfg.SetSize else 1)
    v, e = PIR.NonePrivateQuery(cfg.ChunkSize * cfg.SetSize + 1)
    assert e == 1 and (v == 0).all()
    # same query twice: second one is served from the local cache (pir.go:381-383)
    a, _ = PIR.Query(123, True)
    n1 = PIR.client("FinishedQueryNum")
    b, _ = PIR.Query(123, True)
    assert (a == b).all() and (a == rawDB[123]).all() and PIR.client("FinishedQueryNum") == n1
    with pytest.raises(ValueError):
        pianopir.NewPianoPIR(DBSize + 1, E * 8, rawDB, 8)    # len(rawDB) mismatch (pir.go:483-485)


def test_pir_budget_exhaustion_repreprocesses(oracle):
    """PianoPIR.Query re-runs Preprocessing when FinishedQueryNum == MaxQueryNum (pir.go:527-530)."""
    from pacmann_b200 import pianopir
    DBSize, E = 600, 4
    rawDB = splitmix_db(DBSize, E, seed=24)
    PIR = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 20)
    PIR.SetSeeds(3, 0, 4)
    PIR.Preprocessing()
    maxq = PIR.client("MaxQueryNum")
    rng = np.random.default_rng(25)
    done, ok = 0, 0
    idxs = rng.permutation(DBSize)
    for idx in idxs:                      # distinct indices so the cache never short-circuits
        if PIR.client("FinishedQueryNum") == maxq:
            PIR.SetSeeds(3, 1, 4)         # next epoch's key for the automatic re-preprocessing
        q, err = PIR.Query(int(idx), True)
        done += 1
        if err == 0:
            assert (q == rawDB[idx]).all()
            ok += 1
        else:
            assert (q == 0).all() and err in (pianopir.ERR_TOO_MANY_IN_CHUNK, pianopir.ERR_NO_HIT_HINT)
        if done > maxq + 20:
            break
    assert done > maxq and ok > 0.9 * done
    assert PIR.client("FinishedQueryNum") < maxq     # counter was reset by the re-preprocessing


def test_batch_pir_basic(oracle):
    """TestBatchPIRBasic (pir_test.go:60-202), N = 10^6 x 128 B, BatchSize 32, FailureProbLog2 20."""
    from pacmann_b200 import pianopir
    DBSize, DBEntrySize, BatchSize = 1000000, 16, 32
    rawDB = np.repeat(np.arange(DBSize, dtype=np.uint64)[:, None], DBEntrySize, axis=1)    # rawDB[i][*] = i
    PIR = pianopir.NewSimpleBatchPianoPIR(DBSize, DBEntrySize * 8, BatchSize, rawDB, 20)
    config = PIR.Config()
    assert (config.PartitionNum, config.PartitionSize) == (16, 62500)
    PIR.SetSeeds(31, 32)
    PIR.Preprocessing()
    o_pir = oracle.SimpleBatchPianoPIR(DBSize, DBEntrySize * 8, BatchSize, rawDB.reshape(-1), 20)
    o_pir.preprocessing(key_seed=31, repl_seed=32, threads=8)
    for i in (0, 7, 15):
        assert_same_state(PIR.subPIR(i), o_pir.sub(i))

    rng = np.random.default_rng(33)
    QueryPerPartition = pianopir.QueryPerPartition

    def check_batch(batchQuery, expect_correct):
        responses, err = PIR.Query(batchQuery)
        assert err is None
        o_resp = o_pir.query(batchQuery)
        assert (responses == o_resp).all()            # bit-identical to the oracle, failures included
        for i, idx in enumerate(batchQuery):
            if expect_correct(i):
                assert (responses[i] == rawDB[idx]).all(), f"query[{idx}]"
            else:
                assert (responses[i] == 0).all(), f"query[{idx}] want 0"

    # 1 query per partition: all correct (pir_test.go:87-119)
    batch = [i * config.PartitionSize + int(rng.integers(0, min((i + 1) * config.PartitionSize, DBSize) - i * config.PartitionSize))
             for i in range(config.PartitionNum) for _ in range(QueryPerPartition - 1)]
    check_batch(np.array(batch, np.uint64), lambda i: True)
    # 4 queries per partition: all correct (pir_test.go:121-149)
    batch = [i * config.PartitionSize + int(rng.integers(0, config.PartitionSize)) for i in range(config.PartitionNum) for _ in range(4)]
    check_batch(np.array(batch, np.uint64), lambda i: True)
    # 32 distinct queries, all in partition 0: first QueryPerPartition correct, the rest all-zero (pir_test.go:153-201)
    batch = rng.choice(config.PartitionSize, BatchSize, replace=False).astype(np.uint64)
    check_batch(batch, lambda i: i < QueryPerPartition)
    assert PIR.FinishedBatchNum == o_pir.finished_batch_num and PIR.QueriesMadeInPartition == o_pir.queries_made_in_partition
    for i in (0, 3):
        assert_same_state(PIR.subPIR(i), o_pir.sub(i))
    # every sub-query of a call went to the GPU in one launch
    assert PIR.serverLaunches == 3 and PIR.serverQueries == 16 * (1 + 4 + 2)


def test_batch_pir_many_batches_vs_oracle(oracle):
    """TestBatchPIRPerf's check at a small scale (pir_test.go:242-262): response[0] is zero or correct; here every
    response of 60 random batches is also bit-identical to the oracle's, duplicates and drops included."""
    from pacmann_b200 import pianopir
    DBSize, E, BatchSize = 50000, 112, 32
    rawDB = splitmix_db(DBSize, E, seed=41)
    PIR = pianopir.NewSimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB, 8)
    PIR.SetSeeds(42, 43)
    PIR.Preprocessing()
    o_pir = oracle.SimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB.reshape(-1), 8)
    o_pir.preprocessing(42, 43, threads=8)
    rng = np.random.default_rng(44)
    good = 0
    for it in range(60):
        n = 96 if it % 2 else 32
        batch = rng.integers(0, DBSize, n).astype(np.uint64)
        if it % 5 == 0:
            batch[1] = batch[0]           # duplicate index in one call
        resp, _ = PIR.Query(batch)
        assert (resp == o_pir.query(batch)).all(), f"batch {it}"
        for j in range(n):
            assert (resp[j] == 0).all() or (resp[j] == rawDB[batch[j]]).all()
            good += int((resp[j] == rawDB[batch[j]]).all())
    assert good > 0.5 * 60 * 64
    assert abs(PIR.LocalStorageSize() - o_pir.local_storage_size()) == 0
    assert PIR.CommCostPerBatchOnline() == o_pir.comm_cost_per_batch_online()


def test_batch_pir_redo_preprocessing_when_budget_is_spent(oracle):
    """batch-pir.go:239-245: the batch object re-preprocesses when QueriesMadeInPartition >= MaxQueryNum - 2."""
    from pacmann_b200 import pianopir
    DBSize, E, BatchSize = 1600, 4, 8            # 4 partitions of 400 rows: MaxQueryNum = 119
    rawDB = splitmix_db(DBSize, E, seed=51)
    PIR = pianopir.NewSimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB, 8)
    PIR.SetSeeds(52, 53)
    PIR.Preprocessing()
    o_pir = oracle.SimpleBatchPianoPIR(DBSize, E * 8, BatchSize, rawDB.reshape(-1), 8)
    o_pir.preprocessing(52, 53)
    maxq = PIR.subPIR(0).client("MaxQueryNum")
    rng = np.random.default_rng(54)
    redone = False
    for it in range(maxq + 10):
        batch = rng.integers(0, DBSize, 8).astype(np.uint64)
        before = PIR.QueriesMadeInPartition
        resp, _ = PIR.Query(batch)
        assert (resp == o_pir.query(batch)).all(), f"batch {it}"
        if PIR.QueriesMadeInPartition < before:
            redone = True
        assert PIR.QueriesMadeInPartition == o_pir.queries_made_in_partition
    assert redone
