"""CPU tests: the oracle against every known answer the reference and the standards offer for this path
(SURVEY.md 8c).  Runs without a GPU."""
import json
import os

import numpy as np
import pytest

from util import splitmix_db

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = bytes(range(16))


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(GOLD, "aes_prf_kat.json")))


@pytest.fixture(scope="module")
def dkat():
    return json.load(open(os.path.join(GOLD, "distance_kat.json")))


@pytest.fixture(params=[0, 1], ids=["native", "portable"])
def paths(request, oracle):
    """run each test on the intrinsic path (AES-NI / AVX) and on the portable C path"""
    oracle.lib().orc_force_portable(request.param)
    yield oracle
    oracle.lib().orc_force_portable(0)


def test_fips197_block_vectors(paths, kat):
    for v in kat["fips197"]:
        rk = paths.expand_key(bytes.fromhex(v["key"]))
        assert paths.encrypt_aes128(rk, bytes.fromhex(v["pt"])).hex() == v["ct"]


def test_key_schedule_words(paths, kat):
    for v in kat["schedule"]:
        assert paths.expand_key(bytes.fromhex(v["key"])).tolist() == v["words"]


def test_prf_kat(paths, kat):
    cache = {}
    for v in kat["prf"]:
        rk = cache.setdefault(v["key"], paths.expand_key(bytes.fromhex(v["key"])))
        assert paths.prf(rk, v["tag"], v["x"]) == v["out"], v


def test_survey_prf_table(oracle):
    # SURVEY.md section 4 KAT table (key 000102..0f)
    rk = oracle.expand_key(KEY)
    table = {(0, 0): 0x825b8f87373ba1c6, (1, 0): 0x1e204dad5cdbf50e, (0, 1): 0xa0877cdd63d37ce2, (5, 7): 0x821e9920390baced,
             (104359, 511): 0x3172fbc17501a8d1, (2**29 - 1, 2**35 - 1): 0x69eda142b7162bda}
    for (tag, x), want in table.items():
        assert oracle.prf(rk, tag, x) == want
    assert oracle.prf(rk, 0, 0) & 2047 == 454 and oracle.prf(rk, 1, 0) & 1023 == 270


def test_mmo_is_encrypt_xor_input(paths):
    rk = paths.expand_key(KEY)
    blk = bytes(range(16, 32))
    enc = paths.encrypt_aes128(rk, blk)
    assert paths.aes128_mmo(rk, blk) == bytes(a ^ b for a, b in zip(enc, blk))


def test_xor_slices_reference_constants(paths):
    # TestXORPerf (pir_test.go:279-290)
    p, q = np.full(8, 12312312, np.uint64), np.full(8, 12312, np.uint64)
    paths.xor_slices(p, q)
    assert (p == (12312312 ^ 12312)).all()
    # tail: count = len(src) >> 2 blocks of 4 (aes_amd64.s:136-139)
    a, b = np.arange(7, dtype=np.uint64), np.full(7, 0xFF, np.uint64)
    paths.xor_slices(a, b)
    assert a.tolist() == [0 ^ 0xFF, 1 ^ 0xFF, 2 ^ 0xFF, 3 ^ 0xFF, 4, 5, 6]


def test_l2_golden_bits(paths, dkat):
    for v in dkat["l2"]:
        a = np.array(v["a"], np.uint32).view(np.float32)
        b = np.array(v["b"], np.uint32).view(np.float32)
        assert int(paths.l2dist(a, b).view(np.uint32)) == v["out_bits"], v["dim"]


def test_l2_distance_tolerance_as_reference(oracle):
    # TestDistance (graphann_test.go:15-58): |L2Dist - L2DistSIMD| < 1e-4 at dim 128
    rng = np.random.default_rng(1)
    for _ in range(100):
        v1, v2 = rng.random(128, dtype=np.float32), np.zeros(128, np.float32)
        assert abs(float(oracle.l2dist(v1, v2)) - float(np.sum((v1.astype(np.float64)) ** 2))) < 1e-3


def test_inner_product_matches_scalar(paths, dkat):
    # TestInnerProduct (graphann_test.go:225-247): SIMD == wrapping scalar loop
    for v in dkat["ip"]:
        assert paths.inner_product(np.array(v["a"], np.uint32), np.array(v["b"], np.uint32)) == v["out"]
    # closed form of the scan with v[i][j] = i + j, q[j] = j (graphann_test.go:258-273)
    n, d = 3000, 128
    rows = (np.arange(n, dtype=np.uint64)[:, None] + np.arange(d, dtype=np.uint64)).astype(np.uint32)
    q = np.arange(d, dtype=np.uint32)
    j = np.arange(d, dtype=object)
    closed = sum(int(((i + j) * j).sum()) for i in range(n)) % 2**32
    assert int(paths.ip_scan(rows, q)[0]) == closed
    assert paths.inner_product(np.ones(24, np.uint32), np.ones(24, np.uint32)) == 0   # n % 16 != 0 is rejected


def test_parameter_derivation_and_report_identities(oracle):
    # SURVEY.md 8 derived sizes, and the published report numbers they reproduce exactly
    assert oracle.client_params(2**20, 40) == dict(chunk_size=2048, set_size=512, max_query_num=14195, primary_hint_num=59392, max_query_per_chunk=88)
    assert oracle.client_params(62500, 8) == dict(chunk_size=512, set_size=124, max_query_num=2760, primary_hint_num=3584, max_query_per_chunk=72)
    assert oracle.client_params(200114, 8) == dict(chunk_size=1024, set_size=196, max_query_num=5460, primary_hint_num=7168, max_query_per_chunk=88)
    assert oracle.client_params(18750, 40) == dict(chunk_size=512, set_size=40, max_query_num=1347, primary_hint_num=14848, max_query_per_chunk=104)
    assert oracle.client_params(62500, 20)["primary_hint_num"] == 7680
    sift = oracle.SimpleBatchPianoPIR(10**6, 640, 32, np.zeros(10**6 * 80, np.uint64), 8)
    assert sift.local_storage_size() / 1024 / 1024 == pytest.approx(212.429688, abs=1e-6)       # private-search-report.txt:13
    assert sift.comm_cost_per_batch_online() * 20 * 3 / 1024 == 2130.0                           # private-search-report.txt:21
    support = sift.sub(0).max_query_num // 2
    assert support // (20 * 3) == 23                                                            # "Window Size: 23" (:9)
    assert int(640e6 / support) * 20 * 3 / 1024 == 27173.906250                                 # offline comm per query (:15), batch-pir.go:116
    marco = oracle.SimpleBatchPianoPIR(3201821, 896, 32, np.zeros(3201821 * 112, np.uint64), 8)
    assert marco.comm_cost_per_batch_online() * 20 * 3 / 1024 == 3150.0                          # reproduction/msmarco/README.md:262-265
    assert (marco.partition_num, marco.partition_size, marco.sub(15).db_size) == (16, 200114, 200111)


def test_pir_basic_logic_over_oracle(oracle):
    # TestPIRBasic (pir_test.go:9-58) at its own size
    N, E = 18750, 4
    db = splitmix_db(N, E, seed=2)
    p = oracle.PianoPIR(N, E * 8, db.reshape(-1), 40)
    p.preprocessing(KEY, repl_seed=3)
    rng = np.random.default_rng(4)
    for _ in range(p.max_query_num):
        idx = int(rng.integers(0, N))
        r, rc = p.query(idx)
        assert rc == 0 and (r == db[idx]).all()


def test_batch_pir_basic_logic_over_oracle(oracle):
    # TestBatchPIRBasic (pir_test.go:60-202) at 1/10 of its size (same partition structure, same assertions)
    N, E, B = 100000, 16, 32
    db = np.repeat(np.arange(N, dtype=np.uint64)[:, None], E, axis=1)
    p = oracle.SimpleBatchPianoPIR(N, E * 8, B, db.reshape(-1), 20)
    p.preprocessing(key_seed=5, repl_seed=6, threads=4)
    rng = np.random.default_rng(7)
    ps, pn = p.partition_size, p.partition_num
    one = np.array([i * ps + int(rng.integers(0, ps)) for i in range(pn)], np.uint64)
    assert (p.query(one) == db[one]).all()
    four = np.array([i * ps + int(rng.integers(0, ps)) for i in range(pn) for _ in range(4)], np.uint64)
    assert (p.query(four) == db[four]).all()
    crowd = rng.choice(ps, B, replace=False).astype(np.uint64)
    r = p.query(crowd)
    assert (r[:2] == db[crowd[:2]]).all() and (r[2:] == 0).all()


def test_hint_parity_definition(oracle):
    """The parity of hint h is the XOR of one PRF-selected row per chunk (pir.go:316-339), zero rows in the
    padding, backup group g skipping chunk g: recomputed here from the PRF alone."""
    N, E = 5000, 6
    db = splitmix_db(N, E, seed=8)
    p = oracle.PianoPIR(N, E * 8, db.reshape(-1), 8)
    p.preprocessing(KEY, repl_seed=9)
    rk = p.long_key()
    C, S, P, M = p.chunk_size, p.set_size, p.primary_hint_num, p.max_query_per_chunk
    prim, back = p.table("primary_parity"), p.table("backup_parity")
    for h in [0, 1, P - 1, P, P + M, P + S * M - 1]:
        skip = -1 if h < P else (h - P) // M
        want = np.zeros(E, np.uint64)
        for c in range(S):
            row = c * C + (oracle.prf(rk, h, c) & (C - 1))
            if c != skip and row < N:
                want[:4] ^= db[row][:4]          # E = 6: words 4,5 are never xored (xorSlices tail)
        got = prim[h] if h < P else back.reshape(-1, E)[h - P]
        assert (got == want).all()
    # replacement bookkeeping (pir.go:345-349)
    ridx, rval = p.table("replacement_idx"), p.table("replacement_val")
    for c in (0, S - 1):
        for j in (0, M - 1):
            i = int(ridx[c, j])
            assert i // C == c
            assert (rval[c, j] == (db[i] if i < N else 0)).all()


def test_threaded_and_ranged_preprocessing_agree(oracle):
    N, E = 9000, 8
    db = splitmix_db(N, E, seed=10)
    a = oracle.PianoPIR(N, E * 8, db.reshape(-1), 8)
    a.preprocessing(KEY, repl_seed=1, threads=1)
    b = oracle.PianoPIR(N, E * 8, db.reshape(-1), 8)
    b.preprocessing(KEY, repl_seed=1, threads=5)
    for t in ("primary_parity", "backup_parity", "replacement_idx", "replacement_val"):
        assert (a.table(t) == b.table(t)).all()
    H = a.primary_hint_num + a.set_size * a.max_query_per_chunk
    c = oracle.PianoPIR(N, E * 8, db.reshape(-1), 8)
    c.preprocessing_range(KEY, H // 3, 2 * H // 3)
    full = np.concatenate([a.table("primary_parity"), a.table("backup_parity").reshape(-1, E)])
    part = np.concatenate([c.table("primary_parity"), c.table("backup_parity").reshape(-1, E)])
    assert (part[H // 3:2 * H // 3] == full[H // 3:2 * H // 3]).all()
    assert (part[:H // 3] == 0).all() and (part[2 * H // 3:] == 0).all()


def test_wire_format_round_trip(oracle):
    # private-search.go:371-397 packing and :418-439 unpacking
    rng = np.random.default_rng(11)
    n, dim, m = 50, 12, 4
    vec = rng.standard_normal((n, dim)).astype(np.float32)
    graph = rng.integers(0, n, (n, m), dtype=np.int32)
    raw = oracle.pack_db(vec, graph).reshape(n, (dim + m) // 2)
    as_bytes = raw.view(np.uint8).reshape(n, -1)
    assert (as_bytes[:, :dim * 4].copy().view(np.float32) == vec).all()
    assert (as_bytes[:, dim * 4:].copy().view(np.uint32) == graph.astype(np.uint32)).all()


def test_search_knn_oracle_finds_neighbours(oracle):
    from test_graphann_gpu import make_dataset
    n, dim, m = 3000, 32, 16
    vec, graph = make_dataset(n, dim, m, 12)
    queries = vec[:30] + np.float32(0.001)
    ret, step = oracle.search_knn_basic(vec, graph, np.arange(int(np.sqrt(n))), queries, 5, 15, 3)
    exact = np.argmin(((vec[None, :, :] - queries[:, None, :]) ** 2).sum(-1), axis=1)
    assert (ret[:, 0] == exact).mean() > 0.5          # the chain graph is navigable; sanity only
    assert ((ret >= -1) & (ret < n)).all() and (step >= -1).all()


def test_robust_prune_restatement_properties(oracle):
    """orc_robust_prune follows build_graph.go:169-236: <= m candidates come back unchanged; otherwise the nearest
    candidate is always kept, every kept id is a candidate, exactly m come back (discarded ones refill), and with
    alpha large enough nothing is pruned, so the m nearest come back in distance order."""
    rng = np.random.default_rng(301)
    n, dim, m = 400, 16, 8
    vec = rng.standard_normal((n, dim)).astype(np.float32)
    for u in range(20):
        cand = rng.choice(n, 30, replace=False)
        cand = cand[cand != u]
        assert (oracle.robust_prune(vec, u, cand[:m], m, 1.2) == cand[:m]).all()
        got = oracle.robust_prune(vec, u, cand, m, 1.2)
        d = np.array([oracle.l2dist(vec[u], vec[c]) for c in cand])
        assert len(got) == m and set(got.tolist()) <= set(cand.tolist())
        assert got[0] == cand[np.argmin(d)]
        loose = oracle.robust_prune(vec, u, cand, m, 1e9)
        assert (loose == cand[np.argsort(d, kind="stable")[:m]]).all()


# ---- an independent restatement of graphann/search.go:87-234 in plain Python (Go's container/heap included), used only to
# pin the C oracle's SearchKNN: two transcriptions of the same Go function, written apart, must agree id for id --------------
def _go_heap_push(h, item):          # container/heap.Push: append + up(len-1); Less = dist < (search.go:95-97)
    h.append(item)
    j = len(h) - 1
    while True:
        i = (j - 1) // 2
        if i == j or j == 0 or not (h[j][0] < h[i][0]):
            break
        h[i], h[j] = h[j], h[i]
        j = i


def _go_heap_pop(h):                 # container/heap.Pop: Swap(0, n), down(0, n), remove last
    n = len(h) - 1
    h[0], h[n] = h[n], h[0]
    i = 0
    while True:
        j1 = 2 * i + 1
        if j1 >= n:
            break
        j = j1
        if j1 + 1 < n and h[j1 + 1][0] < h[j1][0]:
            j = j1 + 1
        if not (h[j][0] < h[i][0]):
            break
        h[i], h[j] = h[j], h[i]
        i = j
    return h.pop()


def _search_knn_py(oracle, vec, graph, start_ids, q, k, max_step, parallel):
    """search.go:114-234, non-private graph (GetVertexInfo returns every requested vertex, in order).  Where Go leaves the
    order open (sort.Sort is not stable, map iteration + sort.Slice) the rules of DESIGN.md 2 apply: start ranking by
    (distance, position), final ranking by (distance, id)."""
    dist = lambda v: float(oracle.l2dist(vec[v], q))
    reach, known, heap = {}, {}, []
    ranked = sorted(range(len(start_ids)), key=lambda p: (dist(int(start_ids[p])), p))
    for p in ranked:
        if len(heap) >= parallel:
            break
        v = int(start_ids[p])
        if v in known:
            continue
        known[v] = True
        _go_heap_push(heap, (dist(v), v))
        reach[v] = 0
    for step in range(max_step):
        batch = []
        for _ in range(parallel):
            assert heap, "the test data keeps the queue non-empty (the random-id branch needs Go's rand stream)"
            _, v = _go_heap_pop(heap)
            batch.extend(int(x) for x in graph[v])
        for v in batch:
            if v in known:
                continue
            if not any(int(x) != 0 for x in graph[v]):
                continue
            known[v] = True
            reach[v] = step
            _go_heap_push(heap, (dist(v), v))
    order = sorted(known, key=lambda v: (dist(v), v))
    ret = [order[i] if i < len(order) else -1 for i in range(k)]
    return ret, [reach[v] if v >= 0 else -1 for v in ret]


@pytest.mark.parametrize("integer,dim,m,par,steps", [(False, 16, 8, 2, 8), (True, 8, 8, 3, 7), (True, 4, 6, 2, 10)])
def test_search_knn_matches_an_independent_restatement(oracle, integer, dim, m, par, steps):
    """integer-valued low-dimensional vectors make most distances tie (and fp32 exact in any order): the explore queue's
    arrangement, not just its minimum, then decides which vertex is explored next"""
    from test_graphann_gpu import make_dataset
    n, k = 700, 12
    vec, graph = make_dataset(n, dim, m, 31 + dim, integer)
    if integer:
        vec = np.floor(vec / 64).astype(np.float32)         # values 0..3
    start = np.random.default_rng(3).choice(n, 26, replace=False)
    queries = vec[np.random.default_rng(4).integers(0, n, 8)] + np.float32(0.0 if integer else 0.05)
    ret, step = oracle.search_knn_basic(vec, graph, start, queries, k, steps, par)
    ties = 0
    for i, q in enumerate(queries):
        p_ret, p_step = _search_knn_py(oracle, vec, graph, start, q, k, steps, par)
        assert list(ret[i]) == p_ret and list(step[i]) == p_step, f"query {i}"
        d = [float(oracle.l2dist(vec[v], q)) for v in p_ret if v >= 0]
        ties += sum(1 for a, b in zip(d, d[1:]) if a == b)
    assert not integer or ties > 0          # the tie rules were really exercised
