"""The launch variants the benchmark actually runs, against the CPU oracle, bit for bit.

At production sizes pm_hintgen picks a cooperative launch with a round barrier, 16-warp CTAs (12-14 with the shared
last round switched off), a stream-K shared last round, serpentine sweeps and PRF rounds hoisted for one varying
chunk-id byte.  None of those is reached by the small parity cases, so every one is forced here through pm_tuning_set()
on one MS-MARCO partition at size (200 114 x 896 B, 24 416 hints: 0.2 s of oracle time), on the whole 16-partition
MS-MARCO call the benchmark times, and on BASELINE configs[0] whole (2^20 x 32 B, F = 40) with its 1000 online queries.
"""
import numpy as np
import pytest

from util import oracle_parities, splitmix_db

pytestmark = pytest.mark.gpu

KEY = bytes(range(16))
KNOBS = ("hg_sync", "hg_warps", "hg_ntab", "hg_tail_split", "hg_serpentine", "hg_xbytes", "hg_d2h_groups", "ans_split")


@pytest.fixture()
def knobs(cabi):
    """set launch knobs for one test, restore the defaults afterwards"""
    saved = {k: cabi.tuning_get(k) for k in KNOBS}

    def set_(**kw):
        for k, v in kw.items():
            cabi.tuning_set(k, v)

    yield set_
    for k, v in saved.items():
        cabi.tuning_set(k, v)


@pytest.fixture(scope="module")
def msmarco_partition(oracle):
    """one sub-PIR of the MS-MARCO-shaped batch: 200 114 rows x 896 B -> chunk 1024, set 196, 7168 + 196*88 hints"""
    n_rows, E = 200114, 112
    rows = splitmix_db(n_rows, E, seed=77)
    pir_o = oracle.PianoPIR(n_rows, E * 8, rows.reshape(-1), 8)
    assert (pir_o.chunk_size, pir_o.set_size, pir_o.primary_hint_num, pir_o.max_query_per_chunk) == (1024, 196, 7168, 88)
    pir_o.preprocessing(KEY, repl_seed=3, threads=8)
    return rows, pir_o, oracle_parities(pir_o)


def _run_dev(cabi, db, pir_o, rk, shards=1, which=None):
    """pm_hintgen_dev into a device buffer, `shards`-way hint-set sharding (one call per shard, as one rank each would
    issue it); returns the [H][E] table (rows of shards not in `which` stay zero)."""
    P, S, M, E = pir_o.primary_hint_num, pir_o.set_size, pir_o.max_query_per_chunk, pir_o.entry_u64
    H = P + S * M
    buf = cabi.buf_alloc(H * E * 8)
    cabi.buf_zero(buf, H * E * 8)
    for r in range(shards):
        if which is not None and r not in which:
            continue
        a, b = H * r // shards, H * (r + 1) // shards
        job = cabi.make_job(0, pir_o.db_size, pir_o.chunk_size, S, rk, a, b - a, P, M, parity_out=buf + a * E * 8)
        cabi.hintgen_dev(db, [job])
    db.sync()
    out = cabi.buf_download(buf, np.zeros((H, E), np.uint64))
    cabi.buf_free(buf)
    return out


@pytest.mark.parametrize("sync,warps,tail,serp,xb", [
    (1, 0, 1, 1, 0),     # what the benchmark runs: barrier, 16 warps, shared last round, serpentine, XB = 1
    (0, 0, 1, 1, 0),     # no barrier
    (1, 14, 0, 0, 0),    # round-1 variant of the 8-GPU run: 14-warp CTAs, whole last round, forward sweeps
    (1, 12, 0, 1, 0),
    (0, 13, 1, 0, 0),
    (1, 16, 1, 1, 2),    # PRF with two varying chunk-id bytes (what a 2^20-row instance uses)
    (1, 15, 1, 1, 4),    # generic PRF rounds
    (-1, 0, -1, 1, 0),   # all automatic
])
def test_msmarco_partition_launch_variants(cabi, msmarco_partition, knobs, sync, warps, tail, serp, xb):
    rows, pir_o, want = msmarco_partition
    knobs(hg_sync=sync, hg_warps=warps, hg_tail_split=tail, hg_serpentine=serp, hg_xbytes=xb)
    db = cabi.DB(rows)
    got = _run_dev(cabi, db, pir_o, cabi.expand_key(KEY))
    assert (got == want).all()
    db.close()


@pytest.mark.parametrize("ntab", [1, 4])
def test_msmarco_partition_table_variants(cabi, msmarco_partition, knobs, ntab):
    rows, pir_o, want = msmarco_partition
    knobs(hg_ntab=ntab, hg_sync=1)
    db = cabi.DB(rows)
    assert (_run_dev(cabi, db, pir_o, cabi.expand_key(KEY)) == want).all()
    db.close()


@pytest.mark.parametrize("tail", [0, 1])
def test_msmarco_partition_one_eighth_shards_dev(cabi, msmarco_partition, knobs, tail):
    """each of 8 ranks' shard of the hints, written through pm_hintgen_dev into one table (the 8-GPU layout)"""
    rows, pir_o, want = msmarco_partition
    knobs(hg_sync=1, hg_tail_split=tail)
    db = cabi.DB(rows)
    got = _run_dev(cabi, db, pir_o, cabi.expand_key(KEY), shards=8)
    assert (got == want).all()
    # a single rank's shard leaves the others untouched (the shared last round XORs into zeroed rows of its own only)
    H = want.shape[0]
    one = _run_dev(cabi, db, pir_o, cabi.expand_key(KEY), shards=8, which={3})
    a, b = H * 3 // 8, H * 4 // 8
    assert (one[a:b] == want[a:b]).all() and (one[:a] == 0).all() and (one[b:] == 0).all()
    db.close()


def test_host_buffer_call_launch_groups(cabi, msmarco_partition, knobs):
    """pm_hintgen (host buffers, D2H overlapped by launch group): 1, 3 and 8 groups over an 8-job call"""
    rows, pir_o, want = msmarco_partition
    P, S, M, E = pir_o.primary_hint_num, pir_o.set_size, pir_o.max_query_per_chunk, pir_o.entry_u64
    H = want.shape[0]
    db = cabi.DB(rows)
    rk = cabi.expand_key(KEY)
    for groups in (1, 3, 8):
        knobs(hg_d2h_groups=groups)
        out = np.zeros((H, E), np.uint64)
        jobs = [cabi.make_job(0, pir_o.db_size, pir_o.chunk_size, S, rk, H * r // 8, H * (r + 1) // 8 - H * r // 8, P, M,
                              parity_out=out[H * r // 8:]) for r in range(8)]
        cabi.hintgen(db, jobs)
        assert (out == want).all(), groups
    db.close()


def test_msmarco_full_batch_as_benchmarked(cabi, oracle):
    """The call bench.py times: 3 201 821 x 896 B, 16 sub-PIRs, 390 656 hints in ONE pm_hintgen_dev (cooperative launch,
    41 full rounds + shared last round), every parity against the oracle."""
    import bench
    parts = bench.partitions(bench.N_ROWS, bench.BATCH)
    E = bench.ENTRY_U64
    host_db = bench.gen_db(bench.N_ROWS, E)
    b_o = oracle.SimpleBatchPianoPIR(bench.N_ROWS, E * 8, bench.BATCH, host_db.reshape(-1), bench.FAIL_LOG2)
    b_o.preprocessing(key_seed=bench.SEED, repl_seed=1, threads=16)
    db = cabi.DB(host_db)
    total = sum(p["hints"] for p in parts)
    buf = cabi.buf_alloc(total * E * 8)
    jobs, off = [], 0
    for i, p in enumerate(parts):
        rk = cabi.expand_key(oracle.derive_key(bench.SEED, 0, len(parts), i))
        jobs.append(cabi.make_job(p["row0"], p["n_rows"], p["chunk"], p["set"], rk, 0, p["hints"], p["primary"], p["mqpc"],
                                  parity_out=buf + off * E * 8))
        off += p["hints"]
    assert cabi.tuning_get("hg_sync") == -1 and cabi.tuning_get("hg_tail_split") != 0
    cabi.hintgen_dev(db, jobs)
    db.sync()
    got = cabi.buf_download(buf, np.zeros((total, E), np.uint64))
    off = 0
    for i, p in enumerate(parts):
        assert (got[off:off + p["hints"]] == oracle_parities(b_o.sub(i))).all(), f"sub-PIR {i}"
        off += p["hints"]
    # partition sharding (8 GPUs: rank g owns sub-PIRs 2g, 2g+1 and only their rows): rank 5's call over its own slice
    g = 5
    lo, hi = parts[2 * g]["row0"], parts[2 * g + 1]["row0"] + parts[2 * g + 1]["n_rows"]
    db.close()
    db = cabi.DB(host_db[lo:hi])
    cabi.buf_zero(buf, total * E * 8)
    jobs, off = [], 0
    for i in (2 * g, 2 * g + 1):
        p = parts[i]
        rk = cabi.expand_key(oracle.derive_key(bench.SEED, 0, len(parts), i))
        jobs.append(cabi.make_job(p["row0"] - lo, p["n_rows"], p["chunk"], p["set"], rk, 0, p["hints"], p["primary"], p["mqpc"],
                                  parity_out=buf + off * E * 8))
        off += p["hints"]
    cabi.hintgen_dev(db, jobs)
    db.sync()
    got = cabi.buf_download(buf, np.zeros((off, E), np.uint64))
    off = 0
    for i in (2 * g, 2 * g + 1):
        assert (got[off:off + parts[i]["hints"]] == oracle_parities(b_o.sub(i))).all(), f"partition-sharded sub-PIR {i}"
        off += parts[i]["hints"]
    cabi.buf_free(buf)
    db.close()


def test_cfg0_pir_test_shape_whole(oracle):
    """BASELINE configs[0]: pianopir pir_test N = 2^20 x 32 B, F = 40 (104 448 hints, 53 M PRF evaluations), offline
    hint generation + 1000 online queries, whole client state against the oracle (pir_test.go:204-232 shape)."""
    from pacmann_b200 import pianopir
    from pacmann_b200.keys import derive_key, mix64
    from test_pianopir_gpu import assert_same_state
    DBSize, E = 1 << 20, 4
    rawDB = splitmix_db(DBSize, E, seed=1)
    PIR = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 40)
    cfg = PIR.Config()
    assert (cfg.ChunkSize, cfg.SetSize, PIR.client("primaryHintNum"), PIR.client("maxQueryPerChunk"), PIR.client("MaxQueryNum")) == \
        (2048, 512, 59392, 88, 14195)
    PIR.SetSeeds(key_seed=11, epoch=0, repl_seed=12)
    PIR.Preprocessing()
    o_pir = oracle.PianoPIR(DBSize, E * 8, rawDB.reshape(-1), 40)
    o_pir.preprocessing(derive_key(11, 0, 1, 0), repl_seed=mix64(12, 0), threads=8)
    assert_same_state(PIR, o_pir)
    rng = np.random.default_rng(2)
    for i in range(1000):
        idx = int(rng.integers(0, DBSize))
        q, err = PIR.Query(idx, True)
        o_q, o_rc = o_pir.client_query(idx, True)
        assert err == o_rc and (q == o_q).all()
        if err == 0:
            assert (q == rawDB[idx]).all()
    assert_same_state(PIR, o_pir)


def test_cfg0_shape_resident_client(oracle):
    """the same instance as a one-partition resident client: P = 59 392 program points do not fit the prepare kernel's
    shared-memory mirror, which must then be bypassed, not refused"""
    import ctypes as C
    from pacmann_b200 import cabi
    DBSize, E = 1 << 20, 4
    rawDB = splitmix_db(DBSize, E, seed=1)
    o_pir = oracle.PianoPIR(DBSize, E * 8, rawDB.reshape(-1), 40)
    o_pir.preprocessing(KEY, repl_seed=5, threads=8)
    db = cabi.DB(rawDB)
    part = np.array([0, DBSize, o_pir.chunk_size, o_pir.set_size, o_pir.primary_hint_num, o_pir.max_query_per_chunk,
                     o_pir.get("max_query_num")], np.uint64)
    h = C.c_void_p()
    cabi.check(cabi.lib().pm_client_create(db.h, part.ctypes.data_as(C.c_void_p), 1, C.byref(h)))
    ids, rk, seed = np.zeros(1, np.uint32), cabi.expand_key(KEY), np.array([5], np.uint64)
    cabi.check(cabi.lib().pm_client_preprocess(h, ids.ctypes.data_as(C.c_void_p), 1, rk.ctypes.data_as(C.c_void_p),
                                               seed.ctypes.data_as(C.c_void_p), 0))
    rng = np.random.default_rng(3)
    qn = 500
    idx = rng.choice(DBSize, qn, replace=False)
    queries = np.zeros(qn, dtype=[("part", np.uint32), ("kind", np.uint32), ("idx", np.uint64), ("ds", np.uint64), ("dc", np.uint64)])
    queries["kind"], queries["idx"] = 1, idx
    out, status = np.zeros((qn, E), np.uint64), np.zeros(qn, np.int32)
    for a in range(0, qn, 100):     # five calls of 100 queries
        cabi.check(cabi.lib().pm_client_query_batch(h, queries[a:a + 100].ctypes.data_as(C.c_void_p), 100,
                                                   out[a:a + 100].ctypes.data_as(C.c_void_p), status[a:a + 100].ctypes.data_as(C.c_void_p)))
    for i in range(qn):
        o_q, o_rc = o_pir.client_query(int(idx[i]), True)
        assert status[i] == o_rc and (out[i] == o_q).all()
    P, S, M = o_pir.primary_hint_num, o_pir.set_size, o_pir.max_query_per_chunk
    for t, name, words in [(0, "primary_short_tag", P), (1, "primary_parity", P * E), (2, "primary_program_point", P),
                           (6, "backup_parity", S * M * E), (7, "query_histogram", S)]:
        got = np.zeros(words, np.uint64)
        cabi.check(cabi.lib().pm_client_download(h, 0, t, got.ctypes.data_as(C.c_void_p), words))
        assert (got == o_pir.table(name).reshape(-1)).all(), name
    cabi.check(cabi.lib().pm_client_destroy(h))
    db.close()


def test_every_preprocessing_draws_a_fresh_key(oracle):
    """re-preprocessing one sub-PIR must not reuse the previous key / tags / replacement indices."""
    from pacmann_b200 import pianopir
    DBSize, E = 4000, 4
    rawDB = splitmix_db(DBSize, E, seed=61)
    PIR = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 8)
    PIR.SetSeeds(7, 0, 8)
    PIR.Preprocessing()
    k1, r1, p1 = PIR.long_key(), PIR.table("replacementIdx").copy(), PIR.table("primaryParity").copy()
    PIR.Preprocessing()
    k2, r2, p2 = PIR.long_key(), PIR.table("replacementIdx").copy(), PIR.table("primaryParity").copy()
    assert (k1 != k2).any() and (r1 != r2).any() and (p1 != p2).any()
    # without injected seeds two clients never share a key (OS CSPRNG), and the batch object's sub-PIRs all differ
    A = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 8)
    B = pianopir.NewPianoPIR(DBSize, E * 8, rawDB, 8)
    A.Preprocessing()
    B.Preprocessing()
    assert (A.long_key() != B.long_key()).any()
    q, err = A.Query(17, True)
    assert err == 0 and (q == rawDB[17]).all()
    batch = pianopir.NewSimpleBatchPianoPIR(DBSize, E * 8, 8, rawDB, 8)
    batch.Preprocessing()
    keys = {bytes(batch.subPIR(i).long_key()) for i in range(4)}
    assert len(keys) == 4


@pytest.mark.parametrize("n_rows,E,tail", [(20000, 16, 1), (200114, 112, 1), (5000, 7, 0)])
def test_offset_side_output_matches_the_prf(cabi, oracle, knobs, n_rows, E, tail):
    """pm_hint_job.offsets_out: the kernel also stores PRF(rk, tag_h, c) & (ChunkSize-1) of every (hint, chunk) it evaluates
    -- the resident client's offset index is built from it -- row-major per hint, rows padded to a multiple of 8."""
    rows = splitmix_db(n_rows, E, seed=5 + E)
    pir_o = oracle.PianoPIR(n_rows, E * 8, rows.reshape(-1), 8)
    pir_o.preprocessing(KEY, repl_seed=2, threads=8)
    want_par = oracle_parities(pir_o)
    P, S, M, C = pir_o.primary_hint_num, pir_o.set_size, pir_o.max_query_per_chunk, pir_o.chunk_size
    H, spad = P + S * M, (S + 7) & ~7
    knobs(hg_tail_split=tail)
    db = cabi.DB(rows)
    rk = cabi.expand_key(KEY)
    par = cabi.buf_alloc(H * E * 8)
    off = cabi.buf_alloc(H * spad * 2)
    cabi.buf_zero(off, H * spad * 2)
    cabi.hintgen_dev(db, [cabi.make_job(0, n_rows, C, S, rk, 0, H, P, M, parity_out=par, offsets_out=off)])
    db.sync()
    got_par = cabi.buf_download(par, np.zeros((H, E), np.uint64))
    got_off = cabi.buf_download(off, np.zeros((H, spad), np.uint16))
    assert (got_par == want_par).all()
    tags = np.repeat(np.arange(H, dtype=np.uint64), S)
    cs = np.tile(np.arange(S, dtype=np.uint64), H)
    want_off = (oracle.prf_batch(rk, tags, cs) & np.uint64(C - 1)).astype(np.uint16).reshape(H, S)
    assert (got_off[:, :S] == want_off).all()
    assert (got_off[:, S:] == 0).all()          # the padding columns are never written
    for p_ in (par, off):
        cabi.buf_free(p_)
    db.close()
