// Host-side mirror of the reference's `graphann` package and PIRGraphInfo adapter.  See graphann.hpp.
#include "graphann.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

namespace graphann {

using pianopir::Mix64;

static void check(int rc, const char *what) {
    if (rc != PM_OK) throw std::runtime_error(std::string(what) + ": " + pm_last_error());
}

float L2Dist(const std::vector<float> &v1, const std::vector<float> &v2, int device) {
    float d = 0;
    check(pm_l2_query(v1.data(), 1, v1.size(), v2.data(), &d, device), "pm_l2_query");
    return d;
}
// distances of many host vectors to one query: one launch (the batched form of the L2Dist call sites)
static void dist_many(const std::vector<const float *> &vecs, int64_t dim, const float *query, int device,
                      std::vector<float> *out) {
    out->assign(vecs.size(), 0.f);
    if (vecs.empty()) return;
    std::vector<float> flat(vecs.size() * (size_t)dim);
    for (size_t i = 0; i < vecs.size(); i++) memcpy(&flat[i * dim], vecs[i], (size_t)dim * 4);
    check(pm_l2_query(flat.data(), vecs.size(), (uint64_t)dim, query, out->data(), device), "pm_l2_query");
}

// ---------------------------------------------------------------------------------------------
int BasicGraphInfo::GetVertexInfo(const std::vector<int64_t> &ids, std::vector<Vertex> *out) {
    out->resize(ids.size());
    for (size_t i = 0; i < ids.size(); i++) {
        Vertex &v = (*out)[i];
        v.Id = ids[i];
        v.Neighbors.assign(Graph + ids[i] * M, Graph + (ids[i] + 1) * M);
        v.Vector.assign(Vectors + ids[i] * Dim, Vectors + (ids[i] + 1) * Dim);
    }
    return 0;
}
int BasicGraphInfo::GetStartVertex(std::vector<Vertex> *out) {
    int64_t targetNum = (int64_t)std::sqrt((double)N);
    std::vector<int64_t> batch(targetNum);
    for (int64_t i = 0; i < targetNum; i++) batch[i] = i;
    return GetVertexInfo(batch, out);
}

// ---------------------------------------------------------------------------------------------
void PackEntry(int64_t dim, int64_t m, const float *vector, const int32_t *neighbors, uint64_t *entry) {
    uint8_t *o = (uint8_t *)entry;  // little-endian host: LE f32 bit patterns then LE u32 ids
    memcpy(o, vector, (size_t)dim * 4);
    for (int64_t j = 0; j < m; j++) {
        uint32_t v = (uint32_t)neighbors[j];
        memcpy(o + dim * 4 + j * 4, &v, 4);
    }
}
void Entry2VectorAndNeighbors(int64_t dim, int64_t m, const uint64_t *entry, std::vector<float> *vector,
                              std::vector<int64_t> *neighbors) {
    const uint8_t *e = (const uint8_t *)entry;
    vector->resize(dim);
    memcpy(vector->data(), e, (size_t)dim * 4);
    neighbors->resize(m);
    for (int64_t j = 0; j < m; j++) {
        uint32_t v;
        memcpy(&v, e + dim * 4 + j * 4, 4);
        (*neighbors)[j] = (int64_t)v;
    }
}

PIRGraphInfo::PIRGraphInfo(int64_t N, int64_t Dim, int64_t M, const int32_t *graph, const float *vectors, bool skipPrep,
                           bool nonPrivate, uint64_t seed, int device)
    : N(N), Dim(Dim), M(M), graph(graph), vectors(vectors), skipPrep(skipPrep), NonPrivateMode(nonPrivate), seed(seed),
      device(device) {}
PIRGraphInfo::~PIRGraphInfo() { delete PIR; }

void PIRGraphInfo::Preprocess() {
    DBEntryByteNum = (uint64_t)(Dim * 4 + M * 4);
    const uint64_t E = DBEntryByteNum / 8;
    DBTotalSize = (uint64_t)N * DBEntryByteNum;
    if (shareDBWith && shareDBWith->PIR) {   // the DB is already packed and resident: only a new client is created
        PIR = new pianopir::SimpleBatchPianoPIR((uint64_t)N, DBEntryByteNum, (uint64_t)M, shareDBWith->PIR->db, 8);
    } else {
        rawDB.assign((uint64_t)N * E, 0);
        for (int64_t i = 0; i < N; i++) PackEntry(Dim, M, vectors + i * Dim, graph + i * M, &rawDB[(uint64_t)i * E]);
        PIR = new pianopir::SimpleBatchPianoPIR((uint64_t)N, DBEntryByteNum, (uint64_t)M, rawDB.data(), rawDB.size(), 8, device);
        std::vector<uint64_t>().swap(rawDB);  // the device copy is the server's DB from here on
    }
    PIR->SetSeeds(Mix64(seed, 1), Mix64(seed, 2));
    if (residentClient && !NonPrivateMode) PIR->EnableResidentClient();
    if (skipPrep) PIR->DummyPreprocessing();
    else PIR->Preprocessing();
}

int GetGraphInfo::GetVertexInfoWithDist(const std::vector<int64_t> &ids, const float *, std::vector<Vertex> *out,
                                        std::vector<float> *dists) {
    dists->assign(ids.size(), std::nanf(""));
    return GetVertexInfo(ids, out);
}

int PIRGraphInfo::GetVertexInfo(const std::vector<int64_t> &ids, std::vector<Vertex> *out) {
    std::vector<float> unused;
    return GetVertexInfoWithDist(ids, nullptr, out, &unused);
}

int PIRGraphInfo::GetVertexInfoWithDist(const std::vector<int64_t> &ids, const float *query, std::vector<Vertex> *out,
                                        std::vector<float> *dists) {
    totalQueryNum += (int64_t)ids.size();
    out->resize(ids.size());
    dists->assign(ids.size(), std::nanf(""));
    if (NonPrivateMode) {
        for (size_t i = 0; i < ids.size(); i++) {
            Vertex &v = (*out)[i];
            v.Id = ids[i];
            v.Vector.assign(vectors + ids[i] * Dim, vectors + (ids[i] + 1) * Dim);
            v.Neighbors.assign(graph + ids[i] * M, graph + (ids[i] + 1) * M);
        }
        return 0;
    }
    const uint64_t E = DBEntryByteNum / 8;
    wsIdx.assign(ids.begin(), ids.end());
    wsResp.resize(ids.size() * E);
    if (PIR->QueryFlat(wsIdx.data(), wsIdx.size(), wsResp.data(), query, (uint64_t)Dim, query ? dists->data() : nullptr) != 0) return -1;
    for (size_t i = 0; i < ids.size(); i++) {
        Vertex &v = (*out)[i];
        v.Id = ids[i];
        Entry2VectorAndNeighbors(Dim, M, &wsResp[i * E], &v.Vector, &v.Neighbors);
        bool correctQ = true;
        for (int64_t j = 0; j < M; j++)
            if (v.Neighbors[j] != (int64_t)(uint32_t)graph[ids[i] * M + j]) { correctQ = false; break; }
        if (correctQ) succQueryNum++;
    }
    return 0;
}

int PIRGraphInfo::GetStartVertex(std::vector<Vertex> *out) {
    int64_t targetNum = (int64_t)std::sqrt((double)N);
    std::vector<uint8_t> added((size_t)N, 0);
    out->resize(targetNum);
    uint64_t ctr = 0;
    const uint64_t s = Mix64(seed, 3);
    for (int64_t i = 0; i < targetNum; i++) {
        int64_t x = (int64_t)(Mix64(s, ctr++) % (uint64_t)N);
        while (added[x]) x = (int64_t)(Mix64(s, ctr++) % (uint64_t)N);
        added[x] = 1;
        Vertex &v = (*out)[i];
        v.Id = x;
        v.Vector.assign(vectors + x * Dim, vectors + (x + 1) * Dim);
        v.Neighbors.assign(graph + x * M, graph + (x + 1) * M);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
void GraphANNFrontend::Preprocess() {
    Graph->Preprocess();
    if (Graph->GetStartVertex(&StartVertices) != 0) throw std::runtime_error("GetStartVertex failed");
    UploadStartVertices();
}
// keep the start vertices' vectors resident on the GPU: rows of dim fp32 viewed as dim/2 uint64 (needs an even dim)
void GraphANNFrontend::UploadStartVertices() {
    if (startDb) { pm_db_destroy(startDb); startDb = nullptr; }
    int64_t n, dim, m;
    Graph->GetMetadata(&n, &dim, &m);
    if (StartVertices.empty() || (dim & 1)) return;
    std::vector<uint64_t> rows(StartVertices.size() * (size_t)(dim / 2));
    for (size_t i = 0; i < StartVertices.size(); i++) memcpy(&rows[i * (size_t)(dim / 2)], StartVertices[i].Vector.data(), (size_t)dim * 4);
    check(pm_db_create(rows.data(), StartVertices.size(), (uint64_t)(dim / 2), Graph->Device(), &startDb), "pm_db_create(start vertices)");
    startIds.resize(StartVertices.size());
    for (size_t i = 0; i < startIds.size(); i++) startIds[i] = (int64_t)i;
}
GraphANNFrontend::~GraphANNFrontend() {
    if (startDb) pm_db_destroy(startDb);
}

namespace {
struct VD { float dist; int64_t id; };
// container/heap's up/down/Push/Pop with Less = dist < dist (search.go:92-111)
struct ExploreQueue {
    std::vector<VD> a;
    void up(int64_t j) {
        for (;;) {
            int64_t i = (j - 1) / 2;
            if (i == j || j <= 0 || !(a[j].dist < a[i].dist)) break;
            std::swap(a[i], a[j]);
            j = i;
        }
    }
    void down(int64_t i0, int64_t n) {
        int64_t i = i0;
        for (;;) {
            int64_t j1 = 2 * i + 1;
            if (j1 >= n || j1 < 0) break;
            int64_t j = j1, j2 = j1 + 1;
            if (j2 < n && a[j2].dist < a[j1].dist) j = j2;
            if (!(a[j].dist < a[i].dist)) break;
            std::swap(a[i], a[j]);
            i = j;
        }
    }
    void Push(VD v) { a.push_back(v); up((int64_t)a.size() - 1); }
    VD Pop() {
        int64_t n = (int64_t)a.size() - 1;
        std::swap(a[0], a[n]);
        down(0, n);
        VD v = a.back();
        a.pop_back();
        return v;
    }
    size_t Len() const { return a.size(); }
};
}  // namespace

// Distances of every start vertex to one query (search.go:131-134).  The start vertices are fixed plaintext copies,
// so their vectors are uploaded once (Preprocess) and each search only sends the query.
void GraphANNFrontend::StartDistances(const float *queryVector, int64_t dim, int device, std::vector<float> *out) {
    out->assign(StartVertices.size(), 0.f);
    if (StartVertices.empty()) return;
    if (startDb) {
        check(pm_l2_batch(startDb, (uint64_t)dim, queryVector, 1, startIds.data(), startIds.size(), out->data()), "pm_l2_batch");
        return;
    }
    std::vector<const float *> ptrs;
    for (auto &v : StartVertices) ptrs.push_back(v.Vector.data());
    dist_many(ptrs, dim, queryVector, device, out);
}

int GraphANNFrontend::SearchKNN(const float *queryVector, int64_t k, int64_t maxStep, int64_t parallel, bool benchmarking,
                                std::vector<int64_t> *ret, std::vector<int64_t> *stepRet) {
    int64_t n, dim, m;
    Graph->GetMetadata(&n, &dim, &m);
    const int device = Graph->Device();
    // knownVertices / reachStep (search.go:117-118) as a slot table: the reference keeps whole Vertex objects in maps, but
    // only ids, neighbour lists, reach steps and distances are read back.  A vertex's distance is evaluated once, when it
    // becomes known, and reused by the final ranking (search.go:212-218 recomputes L2Dist on the same vector: same bits).
    std::unordered_map<int64_t, int32_t> slotOf;
    slotOf.reserve((size_t)(maxStep * parallel * m * 2 + 64));
    std::vector<int64_t> knownId, knownStep, nbrPool;
    std::vector<float> knownDist;
    auto add_known = [&](const Vertex &v, float dist, int64_t step) {
        slotOf[v.Id] = (int32_t)knownId.size();
        knownId.push_back(v.Id);
        knownDist.push_back(dist);
        knownStep.push_back(step);
        nbrPool.insert(nbrPool.end(), v.Neighbors.begin(), v.Neighbors.end());
    };
    ExploreQueue toBeExplored;
    const uint64_t rseed = Mix64(randSeed, queryCounter++);
    uint64_t rctr = 0;
    std::vector<float> dists;

    if (!benchmarking) {  // search.go:129-148
        StartDistances(queryVector, dim, device, &dists);
        std::vector<size_t> order(StartVertices.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return dists[a] < dists[b]; });
        for (size_t i = 0; (int64_t)toBeExplored.Len() < parallel && i < order.size(); i++) {
            const Vertex &v = StartVertices[order[i]];
            if (slotOf.count(v.Id)) continue;
            add_known(v, dists[order[i]], 0);
            toBeExplored.Push({dists[order[i]], v.Id});
        }
    }

    std::vector<Vertex> &queryResults = wsResults;
    std::vector<float> &srcDists = wsSrcDists;
    std::vector<int64_t> batchQ;
    std::vector<size_t> fresh, missing;
    std::vector<const float *> ptrs;
    for (int64_t step = 0; step < maxStep; step++) {  // search.go:150-208
        batchQ.clear();
        for (int64_t rept = 0; rept < parallel; rept++) {
            if (toBeExplored.Len() == 0 || benchmarking) {
                for (int64_t i = 0; i < m; i++) batchQ.push_back((int64_t)(Mix64(rseed, rctr++) % (uint64_t)n));
            } else {
                VD item = toBeExplored.Pop();
                const int64_t *nb = &nbrPool[(size_t)slotOf[item.id] * (size_t)m];
                batchQ.insert(batchQ.end(), nb, nb + m);
            }
        }
        if (Graph->GetVertexInfoWithDist(batchQ, benchmarking ? nullptr : queryVector, &queryResults, &srcDists) != 0) return -1;
        if (benchmarking) continue;
        // newly discovered vertices of this step, in batch order (a repeated id is "already known" by its second occurrence)
        fresh.clear();
        for (size_t i = 0; i < queryResults.size(); i++) {
            const Vertex &v = queryResults[i];
            if (slotOf.count(v.Id)) continue;
            bool dup = false;
            for (size_t f : fresh) dup = dup || queryResults[f].Id == v.Id;
            if (dup) continue;
            bool ok = false;
            for (int64_t nb : v.Neighbors) if (nb != 0) { ok = true; break; }   // all-zero list = failed fetch (search.go:192-199)
            if (ok) fresh.push_back(i);
        }
        // their distances (search.go:204): taken from the vertex source when it computed them behind the fetch, one extra
        // launch only for the ones it did not (e.g. entries served from the local cache)
        dists.assign(fresh.size(), 0.f);
        ptrs.clear();
        missing.clear();
        for (size_t t = 0; t < fresh.size(); t++) {
            const float d = srcDists[fresh[t]];
            if (std::isnan(d)) { missing.push_back(t); ptrs.push_back(queryResults[fresh[t]].Vector.data()); }
            else dists[t] = d;
        }
        if (!missing.empty()) {
            std::vector<float> md;
            dist_many(ptrs, dim, queryVector, device, &md);
            for (size_t j = 0; j < missing.size(); j++) dists[missing[j]] = md[j];
        }
        for (size_t t = 0; t < fresh.size(); t++) {
            const Vertex &v = queryResults[fresh[t]];
            add_known(v, dists[t], step);
            toBeExplored.Push({dists[t], v.Id});
        }
    }

    // search.go:210-233
    std::vector<VD> all(knownId.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = {knownDist[i], knownId[i]};
    std::sort(all.begin(), all.end(), [](const VD &a, const VD &b) { return a.dist < b.dist || (a.dist == b.dist && a.id < b.id); });
    ret->assign(k, -1);
    stepRet->assign(k, -1);
    for (int64_t i = 0; i < k && i < (int64_t)all.size(); i++) {
        (*ret)[i] = all[i].id;
        (*stepRet)[i] = knownStep[(size_t)slotOf[all[i].id]];
    }
    return 0;
}

int GraphANNFrontend::SearchKNNBatch(const float *queryVectors, int64_t nq, int64_t k, int64_t maxStep, int64_t parallel,
                                     bool benchmarking, std::vector<int64_t> *ret, std::vector<int64_t> *stepRet) {
    int64_t n, dim, m;
    Graph->GetMetadata(&n, &dim, &m);
    ret->assign((size_t)(nq * k), -1);
    stepRet->assign((size_t)(nq * k), -1);
    std::vector<int64_t> r, s;
    for (int64_t i = 0; i < nq; i++) {  // search.go:236-245: a plain loop
        if (SearchKNN(queryVectors + i * dim, k, maxStep, parallel, benchmarking, &r, &s) != 0) return -1;
        memcpy(&(*ret)[i * k], r.data(), (size_t)k * 8);
        memcpy(&(*stepRet)[i * k], s.data(), (size_t)k * 8);
    }
    return 0;
}

}  // namespace graphann
