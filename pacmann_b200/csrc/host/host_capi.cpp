// Flat C handles over the C++ host mirror (pianopir:: / graphann::) so the Python tests and benchmarks
// can drive it through ctypes.  Not part of the drop-in boundary (that is include/pacmann_cuda.h).
#include <algorithm>
#include <unordered_map>
#include <cstring>
#include <stdexcept>
#include <string>

#include "graphann.hpp"
#include "pianopir.hpp"

#define PMH extern "C" __attribute__((visibility("default")))

static thread_local std::string t_err;
#define PMH_TRY try {
#define PMH_CATCH(rv)                  \
    }                                  \
    catch (const std::exception &e) {  \
        t_err = e.what();              \
        return rv;                     \
    }

using pianopir::PianoPIR;
using pianopir::SimpleBatchPianoPIR;

PMH const char *pmh_last_error() { return t_err.c_str(); }
PMH uint64_t pmh_mix64(uint64_t seed, uint64_t ctr) { return pianopir::Mix64(seed, ctr); }
PMH void pmh_derive_key(uint64_t seed, uint64_t epoch, uint64_t parts, uint64_t i, uint8_t out[16]) {
    auto k = pianopir::DeriveKey(seed, epoch, parts, i);
    memcpy(out, k.b, 16);
}
PMH uint64_t pmh_prf(const uint32_t *rk, uint64_t tag, uint64_t x) {
    std::vector<uint32_t> v(rk, rk + 44);
    return pianopir::PRFEvalWithLongKeyAndTag(v, tag, x);
}

// ---- standalone PianoPIR (owns its DeviceDB) ----
struct PirBox {
    pianopir::DeviceDB *db;
    PianoPIR *pir;
};
PMH void *pmh_pir_new(uint64_t n, uint64_t entry_bytes, const uint64_t *rawDB, uint64_t fail_log2, int device) {
    PMH_TRY
    auto *b = new PirBox();
    b->db = new pianopir::DeviceDB(rawDB, n, entry_bytes / 8, device);
    b->pir = new PianoPIR(n, entry_bytes, b->db, 0, fail_log2);
    return b;
    PMH_CATCH(nullptr)
}
PMH void pmh_pir_free(void *h) {
    auto *b = (PirBox *)h;
    if (!b) return;
    delete b->pir;
    delete b->db;
    delete b;
}
static PianoPIR *as_pir(void *h, int boxed) { return boxed ? ((PirBox *)h)->pir : (PianoPIR *)h; }
PMH void pmh_pir_set_seeds(void *h, int boxed, uint64_t key_seed, uint64_t epoch, uint64_t repl_seed) {
    auto &c = as_pir(h, boxed)->client;
    c.SetSeeds(key_seed, repl_seed);
    c.keyEpoch = epoch;
}
PMH int pmh_pir_preprocessing(void *h, int boxed) {
    PMH_TRY
    as_pir(h, boxed)->Preprocessing();
    return 0;
    PMH_CATCH(-100)
}
PMH int pmh_pir_dummy_preprocessing(void *h, int boxed) {
    PMH_TRY
    as_pir(h, boxed)->DummyPreprocessing();
    return 0;
    PMH_CATCH(-100)
}
PMH int pmh_pir_query(void *h, int boxed, uint64_t idx, int real, uint64_t *out) {
    PMH_TRY
    std::vector<uint64_t> ret;
    int rc = as_pir(h, boxed)->Query(idx, real != 0, &ret);
    memcpy(out, ret.data(), ret.size() * 8);
    return rc;
    PMH_CATCH(-100)
}
PMH int pmh_pir_private_query(void *h, int boxed, const uint32_t *offsets, uint64_t *out) {
    PMH_TRY
    PianoPIR *p = as_pir(h, boxed);
    std::vector<uint32_t> o(offsets, offsets + p->config.SetSize);
    std::vector<uint64_t> ret;
    p->server.PrivateQuery(o, &ret);
    memcpy(out, ret.data(), ret.size() * 8);
    return 0;
    PMH_CATCH(-100)
}
PMH int pmh_pir_nonprivate_query(void *h, int boxed, uint64_t idx, uint64_t *out) {
    PMH_TRY
    std::vector<uint64_t> ret;
    int rc = as_pir(h, boxed)->server.NonePrivateQuery(idx, &ret);
    memcpy(out, ret.data(), ret.size() * 8);
    return rc;
    PMH_CATCH(-100)
}
PMH uint64_t pmh_pir_get(void *h, int boxed, int what) {
    PianoPIR *p = as_pir(h, boxed);
    switch (what) {
    case 0: return p->config.DBEntrySize;
    case 1: return p->config.DBSize;
    case 2: return p->config.ChunkSize;
    case 3: return p->config.SetSize;
    case 4: return p->client.MaxQueryNum;
    case 5: return p->client.primaryHintNum;
    case 6: return p->client.maxQueryPerChunk;
    case 7: return p->client.FinishedQueryNum;
    case 8: return p->config.ThreadNum;
    case 9: return p->config.FailureProbLog2;
    }
    return 0;
}
PMH const uint64_t *pmh_pir_table(void *h, int boxed, int what) {
    auto &c = as_pir(h, boxed)->client;
    switch (what) {
    case 0: return c.primaryShortTag.data();
    case 1: return c.primaryParity.data();
    case 2: return c.primaryProgramPoint.data();
    case 3: return c.replacementIdx.data();
    case 4: return c.replacementVal.data();
    case 5: return c.backupShortTag.data();
    case 6: return c.backupParity.data();
    case 7: return c.QueryHistogram.data();
    }
    return nullptr;
}
PMH const uint32_t *pmh_pir_long_key(void *h, int boxed) { return as_pir(h, boxed)->client.longKey.data(); }
PMH double pmh_pir_local_storage(void *h, int boxed) { return as_pir(h, boxed)->LocalStorageSize(); }
PMH double pmh_pir_comm_cost(void *h, int boxed) { return as_pir(h, boxed)->CommCostPerQuery(); }

// ---- SimpleBatchPianoPIR ----
PMH void *pmh_batch_new(uint64_t n, uint64_t entry_bytes, uint64_t batch, const uint64_t *rawDB, uint64_t len,
                        uint64_t fail_log2, int device) {
    PMH_TRY
    return new SimpleBatchPianoPIR(n, entry_bytes, batch, rawDB, len, fail_log2, device);
    PMH_CATCH(nullptr)
}
PMH void pmh_batch_free(void *h) { delete (SimpleBatchPianoPIR *)h; }
PMH void pmh_batch_set_seeds(void *h, uint64_t key_seed, uint64_t repl_seed) { ((SimpleBatchPianoPIR *)h)->SetSeeds(key_seed, repl_seed); }
PMH int pmh_batch_preprocessing(void *h) {
    PMH_TRY
    ((SimpleBatchPianoPIR *)h)->Preprocessing();
    return 0;
    PMH_CATCH(-100)
}
PMH int pmh_batch_dummy_preprocessing(void *h) {
    PMH_TRY
    ((SimpleBatchPianoPIR *)h)->DummyPreprocessing();
    return 0;
    PMH_CATCH(-100)
}
PMH int pmh_batch_query(void *h, const uint64_t *idx, uint64_t n, uint64_t *out) {
    PMH_TRY
    auto *b = (SimpleBatchPianoPIR *)h;
    std::vector<uint64_t> v(idx, idx + n);
    std::vector<std::vector<uint64_t>> ret;
    int rc = b->Query(v, &ret);
    if (rc != 0) {
        t_err = "index out of range";
        return rc;
    }
    const uint64_t E = b->config.DBEntrySize;
    for (uint64_t i = 0; i < n; i++) memcpy(out + i * E, ret[i].data(), E * 8);
    return 0;
    PMH_CATCH(-100)
}
PMH void *pmh_batch_sub(void *h, uint64_t i) {
    auto *b = (SimpleBatchPianoPIR *)h;
    try {
        b->SyncTablesFromDevice(i);  // resident mode: refresh the host copy of the tables for inspection
    } catch (const std::exception &e) {
        t_err = e.what();
        return nullptr;
    }
    return b->subPIR[i];
}
PMH int pmh_batch_enable_resident(void *h) {
    PMH_TRY
    ((SimpleBatchPianoPIR *)h)->EnableResidentClient();
    return 0;
    PMH_CATCH(-100)
}
PMH uint64_t pmh_batch_get(void *h, int what) {
    auto *b = (SimpleBatchPianoPIR *)h;
    switch (what) {
    case 0: return b->config.PartitionNum;
    case 1: return b->config.PartitionSize;
    case 2: return b->FinishedBatchNum;
    case 3: return b->QueriesMadeInPartition;
    case 4: return b->SupportBatchNum;
    case 5: return b->serverQueries;
    case 6: return b->serverLaunches;
    case 7: return b->CommCostPerBatchOnline();
    case 8: return b->CommCostPerBatchOffline();
    }
    return 0;
}
PMH double pmh_batch_local_storage(void *h) { return ((SimpleBatchPianoPIR *)h)->LocalStorageSize(); }
PMH double pmh_batch_prep_time(void *h) { return ((SimpleBatchPianoPIR *)h)->PreprocessingTime(); }
PMH double pmh_batch_prep_total(void *h) { return ((SimpleBatchPianoPIR *)h)->preprocessingTotal; }
PMH uint64_t pmh_batch_prep_count(void *h) { return ((SimpleBatchPianoPIR *)h)->preprocessingCount; }

// ---- graphann ----
PMH float pmh_l2dist(const float *a, const float *b, uint64_t dim, int device) {
    PMH_TRY
    std::vector<float> x(a, a + dim), y(b, b + dim);
    return graphann::L2Dist(x, y, device);
    PMH_CATCH(-1.0f)
}
struct FrontBox {
    graphann::GetGraphInfo *g;
    graphann::GraphANNFrontend *f;
};
PMH void *pmh_frontend_basic(int64_t n, int64_t dim, int64_t m, const int32_t *graph, const float *vectors) {
    PMH_TRY
    auto *b = new FrontBox();
    b->g = new graphann::BasicGraphInfo(n, dim, m, graph, vectors);
    b->f = new graphann::GraphANNFrontend(b->g);
    return b;
    PMH_CATCH(nullptr)
}
PMH void *pmh_frontend_pir(int64_t n, int64_t dim, int64_t m, const int32_t *graph, const float *vectors, int skip_prep,
                           int non_private, uint64_t seed, int device, int resident) {
    PMH_TRY
    auto *b = new FrontBox();
    auto *pg = new graphann::PIRGraphInfo(n, dim, m, graph, vectors, skip_prep != 0, non_private != 0, seed, device);
    pg->residentClient = resident != 0;
    b->g = pg;
    b->f = new graphann::GraphANNFrontend(b->g);
    return b;
    PMH_CATCH(nullptr)
}
// one more client (its own keys, hint tables and search state) over the rawDB that `other` already keeps on the GPU
PMH void *pmh_frontend_pir_shared(void *other, uint64_t seed, int resident) {
    PMH_TRY
    auto *og = dynamic_cast<graphann::PIRGraphInfo *>(((FrontBox *)other)->g);
    if (!og || !og->PIR) throw std::runtime_error("pmh_frontend_pir_shared: the other frontend has no preprocessed PIR");
    auto *b = new FrontBox();
    auto *pg = new graphann::PIRGraphInfo(og->N, og->Dim, og->M, og->graph, og->vectors, og->skipPrep, false, seed, og->device);
    pg->residentClient = resident != 0;
    pg->shareDBWith = og;
    b->g = pg;
    b->f = new graphann::GraphANNFrontend(b->g);
    return b;
    PMH_CATCH(nullptr)
}
// Client groups for lock-step search: call pmh_frontend_set_group_lanes(first, L) before preprocessing the first
// client, then create clients 1..L-1 with pmh_frontend_pir_lane(first, seed, lane) and preprocess them.
PMH int pmh_frontend_set_group_lanes(void *h, uint32_t lanes) {
    PMH_TRY
    auto *pg = dynamic_cast<graphann::PIRGraphInfo *>(((FrontBox *)h)->g);
    if (!pg || pg->PIR) throw std::runtime_error("pmh_frontend_set_group_lanes: needs a private frontend that is not preprocessed yet");
    pg->groupLanes = lanes ? lanes : 1;
    return 0;
    PMH_CATCH(-100)
}
PMH void *pmh_frontend_pir_lane(void *first, uint64_t seed, uint32_t lane) {
    PMH_TRY
    auto *og = dynamic_cast<graphann::PIRGraphInfo *>(((FrontBox *)first)->g);
    if (!og || !og->PIR || !og->PIR->resident) throw std::runtime_error("pmh_frontend_pir_lane: the first client has no resident client group");
    auto *b = new FrontBox();
    auto *pg = new graphann::PIRGraphInfo(og->N, og->Dim, og->M, og->graph, og->vectors, og->skipPrep, false, seed, og->device);
    pg->residentClient = true;
    pg->shareDBWith = og;
    pg->laneOf = og;
    pg->lane = lane;
    b->g = pg;
    b->f = new graphann::GraphANNFrontend(b->g);
    return b;
    PMH_CATCH(nullptr)
}
PMH int pmh_search_knn_lockstep(void **handles, int64_t n_lanes, const float *queries, int64_t nq, int64_t k, int64_t max_step,
                                int64_t parallel, int benchmarking, int64_t *ret, int64_t *step_ret) {
    PMH_TRY
    std::vector<graphann::GraphANNFrontend *> lanes;
    for (int64_t i = 0; i < n_lanes; i++) lanes.push_back(((FrontBox *)handles[i])->f);
    std::vector<int64_t> r, s;
    int rc = graphann::SearchKNNLockstep(lanes, queries, nq, k, max_step, parallel, benchmarking != 0, &r, &s);
    if (rc != 0) return rc;
    memcpy(ret, r.data(), r.size() * 8);
    memcpy(step_ret, s.data(), s.size() * 8);
    return 0;
    PMH_CATCH(-100)
}
PMH void pmh_device_search_stats(uint64_t *out) { graphann::DeviceSearchStats(out); }
// robustPrune for a batch of vertices: candidates = [n][k] (every vertex k candidates), out = [n][m] padded with -1,
// out_len[n].  vectors are uploaded for the call (rows of dim fp32; dim must be even).
PMH int pmh_robust_prune_batch(const float *vectors, int64_t n_vectors, int64_t dim, const int64_t *us, int64_t n, const int64_t *candidates,
                               int64_t k, int64_t m, float alpha, int device, int64_t *out, int64_t *out_len) {
    PMH_TRY
    if (dim & 1) throw std::runtime_error("pmh_robust_prune_batch: dim must be even");
    pm_db *db = nullptr;
    if (pm_db_create((const uint64_t *)vectors, (uint64_t)n_vectors, (uint64_t)(dim / 2), device, &db) != PM_OK)
        throw std::runtime_error(std::string("pm_db_create: ") + pm_last_error());
    std::vector<int64_t> u(us, us + n);
    std::vector<std::vector<int64_t>> cand((size_t)n), res;
    for (int64_t i = 0; i < n; i++) cand[(size_t)i].assign(candidates + i * k, candidates + (i + 1) * k);
    try {
        graphann::RobustPruneBatch(db, dim, u, cand, m, alpha, &res);
    } catch (...) {
        pm_db_destroy(db);
        throw;
    }
    pm_db_destroy(db);
    const int64_t w = std::max<int64_t>(m, 1);
    for (int64_t i = 0; i < n; i++) {
        out_len[i] = (int64_t)res[(size_t)i].size();
        for (int64_t j = 0; j < w; j++) out[i * w + j] = j < out_len[i] ? res[(size_t)i][(size_t)j] : -1;
    }
    return 0;
    PMH_CATCH(-100)
}
// Self-test of the host-only building blocks (no GPU involved): the open-addressing map against std::unordered_map
// under random inserts / overwrites / lookups / resets, and the worker pool (every index exactly once over many loops of
// varying size, exception propagation).  Returns 0 or the number of the failed check.
PMH int pmh_selftest(int threads) {
    PMH_TRY
    {
        pianopir::FlatMap m;
        std::unordered_map<uint64_t, uint64_t> ref;
        uint64_t x = 88172645463325252ull;
        auto rnd = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
        for (int round = 0; round < 20; round++) {
            m.reset(round % 3 == 0 ? 0 : 500);
            ref.clear();
            for (int i = 0; i < 4000; i++) {
                const uint64_t k = rnd() % 3000, v = rnd();
                if (i % 3 == 0) {
                    const uint64_t *f = m.find(k);
                    auto it = ref.find(k);
                    if ((f != nullptr) != (it != ref.end()) || (f && *f != it->second)) return 1;
                } else {
                    m.put(k, v);
                    ref[k] = v;
                }
            }
            if (m.size() != ref.size()) return 2;
            for (auto &kv : ref) {
                const uint64_t *f = m.find(kv.first);
                if (!f || *f != kv.second) return 3;
            }
        }
    }
    {
        pianopir::WorkerPool pool(threads);
        std::vector<int> hits;
        for (int loop = 0; loop < 3000; loop++) {
            const size_t n = (size_t)(loop * 7919 % 67);
            hits.assign(n, 0);
            pool.ParallelFor(n, [&](size_t i) { hits[i] += 1 + (int)(loop & 1); });
            for (size_t i = 0; i < n; i++)
                if (hits[i] != 1 + (loop & 1)) return 4;
        }
        bool thrown = false;
        try {
            pool.ParallelFor(40, [&](size_t i) { if (i == 17) throw std::runtime_error("boom"); });
        } catch (const std::exception &e) {
            thrown = std::string(e.what()) == "boom";
        }
        if (!thrown) return 5;
        std::vector<int> again(33, 0);
        pool.ParallelFor(33, [&](size_t i) { again[i] = (int)i; });   // the pool still works after an exception
        for (size_t i = 0; i < 33; i++) if (again[i] != (int)i) return 6;
    }
    return 0;
    PMH_CATCH(-100)
}
PMH void pmh_frontend_free(void *h) {
    auto *b = (FrontBox *)h;
    if (!b) return;
    delete b->f;
    delete b->g;
    delete b;
}
PMH int pmh_frontend_preprocess(void *h) {
    PMH_TRY
    ((FrontBox *)h)->f->Preprocess();
    return 0;
    PMH_CATCH(-100)
}
PMH int64_t pmh_frontend_start_ids(void *h, int64_t *out, int64_t cap) {
    auto &sv = ((FrontBox *)h)->f->StartVertices;
    for (int64_t i = 0; i < (int64_t)sv.size() && i < cap; i++) out[i] = sv[i].Id;
    return (int64_t)sv.size();
}
PMH void pmh_frontend_set_start_ids(void *h, const int64_t *ids, int64_t n) {
    auto *b = (FrontBox *)h;
    std::vector<int64_t> v(ids, ids + n);
    // start vertices are plaintext copies of dataset rows (private-search.go:508-531), whichever ids are chosen
    int64_t N, D, M;
    b->g->GetMetadata(&N, &D, &M);
    if (auto *pg = dynamic_cast<graphann::PIRGraphInfo *>(b->g)) {
        b->f->StartVertices.resize(n);
        for (int64_t i = 0; i < n; i++) {
            auto &vx = b->f->StartVertices[i];
            vx.Id = ids[i];
            vx.Vector.assign(pg->vectors + ids[i] * D, pg->vectors + (ids[i] + 1) * D);
            vx.Neighbors.assign(pg->graph + ids[i] * M, pg->graph + (ids[i] + 1) * M);
        }
    } else {
        b->g->GetVertexInfo(v, &b->f->StartVertices);
    }
    try {
        b->f->UploadStartVertices();
    } catch (const std::exception &e) {
        t_err = e.what();
    }
}
PMH void pmh_frontend_set_rand_seed(void *h, uint64_t seed) {
    ((FrontBox *)h)->f->randSeed = seed;
    ((FrontBox *)h)->f->queryCounter = 0;
}
PMH int pmh_frontend_search_knn(void *h, const float *queries, int64_t nq, int64_t k, int64_t max_step, int64_t parallel,
                                int benchmarking, int64_t *ret, int64_t *step_ret) {
    PMH_TRY
    std::vector<int64_t> r, s;
    int rc = ((FrontBox *)h)->f->SearchKNNBatch(queries, nq, k, max_step, parallel, benchmarking != 0, &r, &s);
    if (rc != 0) return rc;
    memcpy(ret, r.data(), r.size() * 8);
    memcpy(step_ret, s.data(), s.size() * 8);
    return 0;
    PMH_CATCH(-100)
}
PMH void *pmh_frontend_pir_handle(void *h) {
    auto *pg = dynamic_cast<graphann::PIRGraphInfo *>(((FrontBox *)h)->g);
    return pg ? pg->PIR : nullptr;
}
PMH int64_t pmh_frontend_stat(void *h, int what) {
    auto *pg = dynamic_cast<graphann::PIRGraphInfo *>(((FrontBox *)h)->g);
    if (!pg) return 0;
    return what == 0 ? pg->totalQueryNum : pg->succQueryNum;
}
