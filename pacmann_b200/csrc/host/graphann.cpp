// Host-side mirror of the reference's `graphann` package and PIRGraphInfo adapter.  See graphann.hpp.
#include "graphann.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
    if (seed != 0) PIR->SetSeeds(Mix64(seed, 1), Mix64(seed, 2));   // test hook; otherwise the client keeps its CSPRNG secrets
    if (residentClient && !NonPrivateMode) {
        if (laneOf && laneOf->PIR) PIR->AttachResidentClient(laneOf->PIR, lane);   // a further lane of laneOf's client group
        else PIR->EnableResidentClient(groupLanes);
    }
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
    unpackResponses(ids, out);
    return 0;
}

// GetVertexInfoWithDist of several lanes of one resident client group: ONE QueryFlatGroup for all of them.
// Raw form: entries[l][i] points at the fetched entry of ids[l][i] (wire format of private-search.go:355-439), valid
// until lane l's next fetch -- the search state reads neighbour lists and vectors in place, no Vertex objects are built.
int PIRGraphInfo::FetchGroupRaw(const std::vector<PIRGraphInfo *> &infos, const std::vector<const std::vector<int64_t> *> &ids,
                                const std::vector<const float *> &queries, const std::vector<std::vector<const uint64_t *> *> &entries,
                                const std::vector<std::vector<float> *> &dists) {
    const size_t L = infos.size();
    std::vector<pianopir::SimpleBatchPianoPIR::GroupCall> calls(L);
    for (size_t l = 0; l < L; l++) {
        PIRGraphInfo *g = infos[l];
        const size_t cnt = ids[l]->size();
        g->totalQueryNum += (int64_t)cnt;
        entries[l]->assign(cnt, nullptr);
        dists[l]->assign(cnt, std::nanf(""));
        g->wsIdx.assign(ids[l]->begin(), ids[l]->end());
        calls[l] = {g->PIR, g->wsIdx.data(), cnt, nullptr, queries[l], queries[l] ? dists[l]->data() : nullptr, 0, entries[l]->data()};
    }
    if (pianopir::SimpleBatchPianoPIR::QueryFlatGroup(calls, (uint64_t)infos[0]->Dim) != 0) return -1;
    pianopir::WorkerPool::Local().ParallelFor(L, [&](size_t l) {   // the reference's correctness accounting (private-search.go:480-504)
        PIRGraphInfo *g = infos[l];
        for (size_t i = 0; i < ids[l]->size(); i++) {
            const uint32_t *nb = (const uint32_t *)((const uint8_t *)(*entries[l])[i] + g->Dim * 4);
            const int32_t *want = g->graph + (*ids[l])[i] * g->M;
            bool correctQ = true;
            for (int64_t j = 0; j < g->M; j++)
                if (nb[j] != (uint32_t)want[j]) { correctQ = false; break; }
            if (correctQ) g->succQueryNum++;
        }
    });
    return 0;
}

// Entry2VectorAndNeighbors of every fetched entry + the reference's correctness accounting (private-search.go:480-504)
void PIRGraphInfo::unpackResponses(const std::vector<int64_t> &ids, std::vector<Vertex> *out) {
    const uint64_t E = DBEntryByteNum / 8;
    for (size_t i = 0; i < ids.size(); i++) {
        Vertex &v = (*out)[i];
        v.Id = ids[i];
        Entry2VectorAndNeighbors(Dim, M, &wsResp[i * E], &v.Vector, &v.Neighbors);
        bool correctQ = true;
        for (int64_t j = 0; j < M; j++)
            if (v.Neighbors[j] != (int64_t)(uint32_t)graph[ids[i] * M + j]) { correctQ = false; break; }
        if (correctQ) succQueryNum++;
    }
}

int PIRGraphInfo::GetStartVertex(std::vector<Vertex> *out) {
    int64_t targetNum = (int64_t)std::sqrt((double)N);
    std::vector<uint8_t> added((size_t)N, 0);
    out->resize(targetNum);
    uint64_t ctr = 0;
    const uint64_t s = seed != 0 ? Mix64(seed, 3) : pianopir::SecureRandom64();
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
    startVersion++;
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
    if (groupStartDb) pm_db_destroy(groupStartDb);
}

// Lock-step groups: the start vertices of all lanes in ONE resident table (lane l = rows [l*kmax, ...)), so that the
// start distances of a round of queries are one pm_l2_batch.  Owned by the group's first frontend; rebuilt when a lane's
// start vertices change.  Returns [act][stride] distances (lane l's first StartVertices.size() entries), or nullptr
// when the lanes cannot be grouped (odd dimension, no start vertices).
const float *GraphANNFrontend::GroupStartDistances(const std::vector<GraphANNFrontend *> &lanes, const float *queries, int64_t act, int64_t dim,
                                                   size_t *stride) {
    if (dim & 1) return nullptr;
    uint64_t stamp = lanes.size();
    size_t kmax = 0;
    for (auto *f : lanes) {
        stamp = Mix64(stamp, (uint64_t)(uintptr_t)f ^ f->startVersion);
        kmax = std::max(kmax, f->StartVertices.size());
    }
    if (kmax == 0) return nullptr;
    if (!groupStartDb || groupStamp != stamp) {
        if (groupStartDb) { pm_db_destroy(groupStartDb); groupStartDb = nullptr; }
        std::vector<uint64_t> rows(lanes.size() * kmax * (size_t)(dim / 2), 0);
        groupIds.assign(lanes.size() * kmax, -1);
        for (size_t l = 0; l < lanes.size(); l++)
            for (size_t i = 0; i < lanes[l]->StartVertices.size(); i++) {
                memcpy(&rows[(l * kmax + i) * (size_t)(dim / 2)], lanes[l]->StartVertices[i].Vector.data(), (size_t)dim * 4);
                groupIds[l * kmax + i] = (int64_t)(l * kmax + i);
            }
        check(pm_db_create(rows.data(), lanes.size() * kmax, (uint64_t)(dim / 2), Graph->Device(), &groupStartDb), "pm_db_create(group start vertices)");
        groupStamp = stamp;
    }
    groupDists.resize((size_t)act * kmax);
    check(pm_l2_batch(groupStartDb, (uint64_t)dim, queries, (uint64_t)act, groupIds.data(), kmax, groupDists.data()), "pm_l2_batch");
    *stride = kmax;
    return groupDists.data();
}

// container/heap's up/down/Push/Pop with Less = dist < dist (search.go:92-111)
void ExploreQueue::up(int64_t j) {
    for (;;) {
        int64_t i = (j - 1) / 2;
        if (i == j || j <= 0 || !(a[j].dist < a[i].dist)) break;
        std::swap(a[i], a[j]);
        j = i;
    }
}
void ExploreQueue::down(int64_t i0, int64_t n) {
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
void ExploreQueue::Push(VD v) { a.push_back(v); up((int64_t)a.size() - 1); }
VD ExploreQueue::Pop() {
    int64_t n = (int64_t)a.size() - 1;
    std::swap(a[0], a[n]);
    down(0, n);
    VD v = a.back();
    a.pop_back();
    return v;
}

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

// ---- SearchKNN (search.go:114-234) as a resumable state: Begin, then NextBatch / Consume once per step, then Finish ----
void SearchState::addKnown(const Vertex &v, float dist, int64_t step) {
    slotOf.put((uint64_t)v.Id, knownId.size());
    knownId.push_back(v.Id);
    knownDist.push_back(dist);
    knownStep.push_back(step);
    nbrPool.insert(nbrPool.end(), v.Neighbors.begin(), v.Neighbors.end());
}

void SearchState::Begin(GraphANNFrontend *front, const float *q, int64_t k_, int64_t maxStep_, int64_t parallel_, bool benchmarking_,
                        const float *startDists) {
    f = front;
    queryVector = q;
    k = k_; maxStep = maxStep_; parallel = parallel_; benchmarking = benchmarking_;
    f->Graph->GetMetadata(&n, &dim, &m);
    device = f->Graph->Device();
    // knownVertices / reachStep (search.go:117-118) as a slot table: the reference keeps whole Vertex objects in maps, but
    // only ids, neighbour lists, reach steps and distances are read back.  A vertex's distance is evaluated once, when it
    // becomes known, and reused by the final ranking (search.go:212-218 recomputes L2Dist on the same vector: same bits).
    slotOf.reset((size_t)(maxStep * parallel * m + 64));
    knownId.clear(); knownStep.clear(); nbrPool.clear(); knownDist.clear();
    toBeExplored.a.clear();
    rseed = Mix64(f->randSeed, f->queryCounter++);
    rctr = 0;
    step = 0;
    if (!benchmarking) {  // search.go:129-148
        if (startDists) dists.assign(startDists, startDists + f->StartVertices.size());
        else f->StartDistances(queryVector, dim, device, &dists);
        // The reference sorts all start vertices by distance (stable here) and takes the first `parallel` distinct ones
        // (search.go:135-148).  Only that prefix is read, so a partial sort under the same total order (distance, then
        // input position) gives the same vertices; the full sort only runs if duplicates exhaust the prefix.
        std::vector<size_t> &order = startOrder;
        order.resize(f->StartVertices.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        auto before = [&](size_t a, size_t b) { return dists[a] < dists[b] || (dists[a] == dists[b] && a < b); };
        size_t sorted = std::min(order.size(), (size_t)parallel * 4 + 8);
        std::partial_sort(order.begin(), order.begin() + (long)sorted, order.end(), before);
        for (size_t i = 0; (int64_t)toBeExplored.Len() < parallel && i < order.size(); i++) {
            if (i == sorted) {   // more duplicates than the sorted prefix: finish the sort
                std::sort(order.begin() + (long)sorted, order.end(), before);
                sorted = order.size();
            }
            const Vertex &v = f->StartVertices[order[i]];
            if (slotOf.has((uint64_t)v.Id)) continue;
            addKnown(v, dists[order[i]], 0);
            toBeExplored.Push({dists[order[i]], v.Id});
        }
    }
}

// the ids this step fetches (search.go:150-167); false once maxStep steps have been made
bool SearchState::NextBatch(std::vector<int64_t> *batchQ) {
    if (step >= maxStep) return false;
    batchQ->clear();
    for (int64_t rept = 0; rept < parallel; rept++) {
        if (toBeExplored.Len() == 0 || benchmarking) {
            for (int64_t i = 0; i < m; i++) batchQ->push_back((int64_t)(Mix64(rseed, rctr++) % (uint64_t)n));
        } else {
            VD item = toBeExplored.Pop();
            const int64_t *nb = &nbrPool[(size_t)*slotOf.find((uint64_t)item.id) * (size_t)m];
            batchQ->insert(batchQ->end(), nb, nb + m);
        }
    }
    return true;
}

// The fetched vertices of a step as the search reads them: Vertex objects (GetGraphInfo interface) or entries in the
// wire format of private-search.go:355-439 read in place.
struct VertexAccess {
    const std::vector<Vertex> &v;
    size_t size() const { return v.size(); }
    int64_t id(size_t i) const { return v[i].Id; }
    int64_t neighbor(size_t i, int64_t j) const { return v[i].Neighbors[(size_t)j]; }
    int64_t degree(size_t i) const { return (int64_t)v[i].Neighbors.size(); }
    const float *vector(size_t i) const { return v[i].Vector.data(); }
};
struct RawAccess {
    const std::vector<int64_t> &ids;
    const std::vector<const uint64_t *> &e;
    int64_t dim, m;
    size_t size() const { return ids.size(); }
    int64_t id(size_t i) const { return ids[i]; }
    int64_t neighbor(size_t i, int64_t j) const {
        uint32_t v;
        memcpy(&v, (const uint8_t *)e[i] + dim * 4 + j * 4, 4);
        return (int64_t)v;
    }
    int64_t degree(size_t) const { return m; }
    const float *vector(size_t i) const { return (const float *)e[i]; }
};

// search.go:169-208 for the fetched vertices of this step, in three parts so that a lock-step driver can evaluate the
// distances the vertex source did not provide for ALL lanes with one launch:
//   collect -> (L2Dist of MissingVectors() to the query) -> apply
template <class A>
void SearchState::collect(const A &res, const std::vector<float> &srcDists) {
    fresh.clear();
    ptrs.clear();
    missing.clear();
    if (benchmarking) return;
    // newly discovered vertices of this step, in batch order (a repeated id is "already known" by its second occurrence)
    for (size_t i = 0; i < res.size(); i++) {
        const int64_t id = res.id(i);
        if (slotOf.has((uint64_t)id)) continue;
        bool dup = false;
        for (size_t t : fresh) dup = dup || res.id(t) == id;
        if (dup) continue;
        bool ok = false;
        for (int64_t j = 0, d = res.degree(i); j < d; j++) if (res.neighbor(i, j) != 0) { ok = true; break; }   // all-zero list = failed fetch (search.go:192-199)
        if (ok) fresh.push_back(i);
    }
    // their distances (search.go:204): taken from the vertex source when it computed them behind the fetch; the others
    // (e.g. entries served from the local cache) are listed for one extra launch
    dists.assign(fresh.size(), 0.f);
    for (size_t t = 0; t < fresh.size(); t++) {
        const float d = srcDists[fresh[t]];
        if (std::isnan(d)) { missing.push_back(t); ptrs.push_back(res.vector(fresh[t])); }
        else dists[t] = d;
    }
}

template <class A>
void SearchState::apply(const A &res, const float *missingDists) {
    const int64_t thisStep = step++;
    if (benchmarking) return;
    for (size_t j = 0; j < missing.size(); j++) dists[missing[j]] = missingDists[j];
    for (size_t t = 0; t < fresh.size(); t++) {
        const size_t i = fresh[t];
        const int64_t id = res.id(i);
        slotOf.put((uint64_t)id, knownId.size());
        knownId.push_back(id);
        knownDist.push_back(dists[t]);
        knownStep.push_back(thisStep);
        for (int64_t j = 0, d = res.degree(i); j < d; j++) nbrPool.push_back(res.neighbor(i, j));
        toBeExplored.Push({dists[t], id});
    }
}

void SearchState::CollectFresh(const std::vector<Vertex> &queryResults, const std::vector<float> &srcDists) {
    collect(VertexAccess{queryResults}, srcDists);
}
void SearchState::ApplyFresh(const std::vector<Vertex> &queryResults, const float *missingDists) {
    apply(VertexAccess{queryResults}, missingDists);
}
void SearchState::CollectFreshRaw(const std::vector<int64_t> &ids, const std::vector<const uint64_t *> &entries, const std::vector<float> &srcDists) {
    collect(RawAccess{ids, entries, dim, m}, srcDists);
}
void SearchState::ApplyFreshRaw(const std::vector<int64_t> &ids, const std::vector<const uint64_t *> &entries, const float *missingDists) {
    apply(RawAccess{ids, entries, dim, m}, missingDists);
}

void SearchState::Consume(const std::vector<Vertex> &queryResults, const std::vector<float> &srcDists) {
    CollectFresh(queryResults, srcDists);
    std::vector<float> md;
    if (!ptrs.empty()) dist_many(ptrs, dim, queryVector, device, &md);
    ApplyFresh(queryResults, md.data());
}

// search.go:210-233
void SearchState::Finish(int64_t *ret, int64_t *stepRet) {
    std::vector<VD> all(knownId.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = {knownDist[i], knownId[i]};
    std::sort(all.begin(), all.end(), [](const VD &a, const VD &b) { return a.dist < b.dist || (a.dist == b.dist && a.id < b.id); });
    for (int64_t i = 0; i < k; i++) { ret[i] = -1; stepRet[i] = -1; }
    for (int64_t i = 0; i < k && i < (int64_t)all.size(); i++) {
        ret[i] = all[i].id;
        stepRet[i] = knownStep[(size_t)*slotOf.find((uint64_t)all[i].id)];
    }
}

int GraphANNFrontend::SearchKNN(const float *queryVector, int64_t k, int64_t maxStep, int64_t parallel, bool benchmarking,
                                std::vector<int64_t> *ret, std::vector<int64_t> *stepRet) {
    SearchState &st = wsState;
    st.Begin(this, queryVector, k, maxStep, parallel, benchmarking);
    std::vector<int64_t> &batchQ = wsBatch;
    while (st.NextBatch(&batchQ)) {
        if (Graph->GetVertexInfoWithDist(batchQ, benchmarking ? nullptr : queryVector, &wsResults, &wsSrcDists) != 0) return -1;
        st.Consume(wsResults, wsSrcDists);
    }
    ret->assign((size_t)k, -1);
    stepRet->assign((size_t)k, -1);
    st.Finish(ret->data(), stepRet->data());
    return 0;
}

int GraphANNFrontend::SearchKNNBatch(const float *queryVectors, int64_t nq, int64_t k, int64_t maxStep, int64_t parallel,
                                     bool benchmarking, std::vector<int64_t> *ret, std::vector<int64_t> *stepRet) {
    int64_t n, dim, m;
    Graph->GetMetadata(&n, &dim, &m);
    ret->assign((size_t)(nq * k), -1);
    stepRet->assign((size_t)(nq * k), -1);
    // a single resident client is a lock-step group of one lane: its searches run with the frontier on the GPU as well
    // (same results; the host only enqueues the steps)
    if (auto *pg = dynamic_cast<PIRGraphInfo *>(Graph))
        if (!pg->NonPrivateMode && pg->PIR && pg->PIR->resident && pg->PIR->clientLanes == 1 && pg->PIR->ownsClient && nq > 0) {
            std::vector<GraphANNFrontend *> one{this};
            return SearchKNNLockstep(one, queryVectors, nq, k, maxStep, parallel, benchmarking, ret, stepRet);
        }
    std::vector<int64_t> r, s;
    for (int64_t i = 0; i < nq; i++) {  // search.go:236-245: a plain loop
        if (SearchKNN(queryVectors + i * dim, k, maxStep, parallel, benchmarking, &r, &s) != 0) return -1;
        memcpy(&(*ret)[i * k], r.data(), (size_t)k * 8);
        memcpy(&(*stepRet)[i * k], s.data(), (size_t)k * 8);
    }
    return 0;
}

// ---- lock-step search with the frontier on the GPU (pm_search_*) ----
// Usable when the lanes are clients of ONE resident pm_client group.  The host only enqueues the steps and keeps the
// counters that need no entry data: the batch budget of every lane (batch-pir.go:239-245: a due Preprocessing() is
// issued between two steps, after the fetch has been applied) and the statistics.  A lane one of whose sub-PIRs could
// reach its query budget within the round (pir.go:527-530 would re-preprocess in the middle of a call) leaves the
// device path for the rest of its batch epoch: its device caches are pulled into the host-side localCache and it
// searches through the host path until the next Preprocessing() of its whole batch.
static bool deviceSearchEnabled() {   // PM_SEARCH_DEVICE=0 keeps the frontier on the host (read per call: tests flip it)
    const char *v = getenv("PM_SEARCH_DEVICE");
    return !(v && *v == '0');
}

static pm_search *deviceSearchFor(const std::vector<GraphANNFrontend *> &lanes, const std::vector<PIRGraphInfo *> &infos, int64_t maxStep,
                                  int64_t parallel) {
    // the lanes must be exactly the clients of one group, lane l in parts [l*PN, (l+1)*PN)
    pianopir::SimpleBatchPianoPIR *owner = nullptr;
    const size_t L = lanes.size();
    for (size_t l = 0; l < L; l++) {
        pianopir::SimpleBatchPianoPIR *p = infos[l] ? infos[l]->PIR : nullptr;
        if (!p || !p->resident || !p->rclient || infos[l]->NonPrivateMode) return nullptr;
        if (p->rclient != infos[0]->PIR->rclient || p->clientLanes != L) return nullptr;
        if (p->partBase != (uint32_t)(l * p->config.PartitionNum)) return nullptr;
        if (lanes[l]->StartVertices.size() != lanes[0]->StartVertices.size() || lanes[l]->StartVertices.empty()) return nullptr;
        if (p->ownsClient) owner = p;
    }
    if (!owner) return nullptr;
    int64_t n, dim, m;
    lanes[0]->Graph->GetMetadata(&n, &dim, &m);
    if ((uint64_t)(parallel * m) < owner->config.PartitionNum) return nullptr;
    uint64_t key = Mix64(Mix64((uint64_t)maxStep, (uint64_t)parallel), L);
    for (auto *f : lanes) key = Mix64(key, (uint64_t)(uintptr_t)f ^ f->startVersion);
    if (owner->devSearch && owner->devSearchKey == key) return owner->devSearch;
    if (owner->devSearch) { pm_search_destroy(owner->devSearch); owner->devSearch = nullptr; }
    pm_search_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n = (uint64_t)n; cfg.dim = (uint64_t)dim; cfg.m = (uint64_t)m;
    cfg.max_step = (uint64_t)maxStep; cfg.parallel = (uint64_t)parallel; cfg.lanes = L;
    cfg.partition_num = owner->config.PartitionNum; cfg.partition_size = owner->config.PartitionSize;
    cfg.n_start = lanes[0]->StartVertices.size();
    cfg.cache_entries = 0;
    for (auto *p : owner->subPIR) cfg.cache_entries = std::max<uint64_t>(cfg.cache_entries, p->client.MaxQueryNum);
    pm_search *s = nullptr;
    if (pm_search_create(owner->rclient, &cfg, &s) != PM_OK) return nullptr;   // e.g. shapes the kernels do not cover: host path
    std::vector<int64_t> ids(cfg.n_start);
    std::vector<float> vec(cfg.n_start * (size_t)dim);
    std::vector<int32_t> nb(cfg.n_start * (size_t)m);
    std::vector<uint64_t> dseed(cfg.partition_num);
    for (size_t l = 0; l < L; l++) {
        for (size_t i = 0; i < cfg.n_start; i++) {
            const Vertex &v = lanes[l]->StartVertices[i];
            ids[i] = v.Id;
            memcpy(&vec[i * (size_t)dim], v.Vector.data(), (size_t)dim * 4);
            for (int64_t j = 0; j < m; j++) nb[i * (size_t)m + (size_t)j] = (int32_t)v.Neighbors[(size_t)j];
        }
        check(pm_search_set_start(s, (uint32_t)l, ids.data(), vec.data(), nb.data()), "pm_search_set_start");
        for (uint64_t p = 0; p < cfg.partition_num; p++) dseed[p] = Mix64(infos[l]->PIR->subPIR[p]->client.dummySeed, 0x5EA7C4);
        check(pm_search_set_dummy_seed(s, (uint32_t)l, dseed.data()), "pm_search_set_dummy_seed");
    }
    owner->devSearch = s;
    owner->devSearchKey = key;
    return s;
}

static std::atomic<uint64_t> g_deviceRounds{0}, g_deviceQueries{0}, g_hostModeQueries{0};
void DeviceSearchStats(uint64_t out[3]) { out[0] = g_deviceRounds.load(); out[1] = g_deviceQueries.load(); out[2] = g_hostModeQueries.load(); }

// one round of the device path: lanes[sel[a]] searches query qidx[a]
static int deviceRound(pm_search *s, const std::vector<GraphANNFrontend *> &lanes, const std::vector<PIRGraphInfo *> &infos,
                       const std::vector<uint32_t> &sel, const float *queryVectors, const std::vector<int64_t> &qidx, int64_t dim, int64_t m,
                       int64_t k, int64_t maxStep, int64_t parallel, bool benchmarking, int64_t *ret, int64_t *stepRet) {
    const size_t act = sel.size();
    const size_t n = (size_t)(parallel * m);
    g_deviceRounds += 1;
    g_deviceQueries += act;
    std::vector<float> q(act * (size_t)dim);
    std::vector<uint64_t> rseeds(act);
    for (size_t a = 0; a < act; a++) {
        memcpy(&q[a * (size_t)dim], queryVectors + qidx[a] * dim, (size_t)dim * 4);
        GraphANNFrontend *f = lanes[sel[a]];
        rseeds[a] = Mix64(f->randSeed, f->queryCounter++);
    }
    check(pm_search_begin(s, sel.data(), act, q.data(), rseeds.data(), (uint64_t)k, benchmarking ? 1 : 0), "pm_search_begin");
    bool applied = true;   // nothing to apply before the first fetch
    for (int64_t step = 0; step < maxStep; step++) {
        check(pm_search_fetch(s, applied ? 0 : 1), "pm_search_fetch");
        applied = false;
        bool anyDue = false;
        std::vector<char> due(act, 0);
        for (size_t a = 0; a < act; a++) {
            due[a] = infos[sel[a]]->PIR->DeviceFetchAccounting(n) ? 1 : 0;
            anyDue = anyDue || due[a];
        }
        if (anyDue) {   // batch-pir.go:239-245: the call's responses are booked first, then the whole batch is preprocessed again
            check(pm_search_apply(s), "pm_search_apply");
            applied = true;
            for (size_t a = 0; a < act; a++)
                if (due[a]) infos[sel[a]]->PIR->Preprocessing();
        }
    }
    std::vector<int64_t> r(act * (size_t)k), st(act * (size_t)k);
    std::vector<uint64_t> stats(act * 3), fin(act * infos[0]->PIR->config.PartitionNum);
    check(pm_search_finish(s, applied ? 0 : 1, r.data(), st.data(), stats.data(), fin.data()), "pm_search_finish");
    const uint64_t PN = infos[0]->PIR->config.PartitionNum;
    for (size_t a = 0; a < act; a++) {
        memcpy(ret + qidx[a] * k, &r[a * (size_t)k], (size_t)k * 8);
        memcpy(stepRet + qidx[a] * k, &st[a * (size_t)k], (size_t)k * 8);
        PIRGraphInfo *g = infos[sel[a]];
        g->totalQueryNum += (int64_t)stats[a * 3];
        g->succQueryNum += (int64_t)stats[a * 3 + 1];
        g->PIR->AbsorbDeviceRound(&fin[a * PN], stats[a * 3 + 2]);
    }
    return 0;
}

// Lock-step search over several lanes.  Query i goes to lane i % L; every lane runs its queries in
// order exactly as its own SearchKNNBatch would -- same client state, same results -- but each step fetches the
// vertices of all lanes together: one FetchGroupRaw (one device call) per step instead of one per lane.
int SearchKNNLockstep(const std::vector<GraphANNFrontend *> &lanes, const float *queryVectors, int64_t nq, int64_t k, int64_t maxStep,
                      int64_t parallel, bool benchmarking, std::vector<int64_t> *ret, std::vector<int64_t> *stepRet) {
    const int64_t L = (int64_t)lanes.size();
    if (L == 0) return -1;
    int64_t n, dim, m;
    lanes[0]->Graph->GetMetadata(&n, &dim, &m);
    ret->assign((size_t)(nq * k), -1);
    stepRet->assign((size_t)(nq * k), -1);
    std::vector<PIRGraphInfo *> infos((size_t)L);
    bool groupable = true;
    for (int64_t l = 0; l < L; l++) {
        infos[(size_t)l] = dynamic_cast<PIRGraphInfo *>(lanes[(size_t)l]->Graph);
        groupable = groupable && infos[(size_t)l] != nullptr && !infos[(size_t)l]->NonPrivateMode;
    }
    if (groupable && deviceSearchEnabled()) {
        if (pm_search *ds = deviceSearchFor(lanes, infos, maxStep, parallel)) {
            for (int64_t base = 0; base < nq; base += L) {
                const int64_t act = std::min(L, nq - base);
                std::vector<uint32_t> sel;
                std::vector<int64_t> qidx;
                std::vector<int64_t> hostLanes;
                for (int64_t l = 0; l < act; l++) {
                    pianopir::SimpleBatchPianoPIR *p = infos[(size_t)l]->PIR;
                    if (!p->hostMode && !p->DeviceRoundIsSafe((uint64_t)maxStep, (size_t)(parallel * m))) p->EnterHostMode(ds);
                    if (p->hostMode) hostLanes.push_back(l);
                    else { sel.push_back((uint32_t)l); qidx.push_back(base + l); }
                }
                if (!sel.empty() &&
                    deviceRound(ds, lanes, infos, sel, queryVectors, qidx, dim, m, k, maxStep, parallel, benchmarking, ret->data(), stepRet->data()) != 0)
                    return -1;
                std::vector<int64_t> r1, s1;
                g_hostModeQueries += hostLanes.size();
                for (int64_t l : hostLanes) {   // rare: this lane's budget epoch is about to end, it searches through the host path
                    if (lanes[(size_t)l]->SearchKNN(queryVectors + (base + l) * dim, k, maxStep, parallel, benchmarking, &r1, &s1) != 0) return -1;
                    memcpy(&(*ret)[(size_t)((base + l) * k)], r1.data(), (size_t)k * 8);
                    memcpy(&(*stepRet)[(size_t)((base + l) * k)], s1.data(), (size_t)k * 8);
                }
            }
            return 0;
        }
    }
    std::vector<std::vector<int64_t>> batch((size_t)L);
    std::vector<std::vector<Vertex>> results((size_t)L);
    std::vector<std::vector<const uint64_t *>> rawEntries((size_t)L);
    std::vector<std::vector<float>> srcDists((size_t)L);
    std::vector<const float *> qptr((size_t)L);
    std::vector<char> more((size_t)L);
    std::vector<size_t> missBase((size_t)L);
    std::vector<float> missA, missB, missDist;
    // PM_HOST_PROFILE=1: where a lock-step step spends its time (printed once per call)
    static const bool prof = getenv("PM_HOST_PROFILE") != nullptr;
    double tBegin = 0, tNext = 0, tFetch = 0, tCollect = 0, tMiss = 0, tApply = 0;
    uint64_t nSteps = 0, nMissTotal = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    for (int64_t base = 0; base < nq; base += L) {
        const int64_t act = std::min(L, nq - base);
        auto t0 = now();
        // start-vertex distances of all lanes in one launch (search.go:131-134 per lane)
        const float *groupDists = nullptr;
        size_t groupStride = 0;
        if (!benchmarking) groupDists = lanes[0]->GroupStartDistances(lanes, queryVectors + base * dim, act, dim, &groupStride);
        pianopir::WorkerPool &pool = pianopir::WorkerPool::Local();
        pool.ParallelFor((size_t)act, [&](size_t l) {
            qptr[l] = queryVectors + (base + (int64_t)l) * dim;
            lanes[l]->wsState.Begin(lanes[l], qptr[l], k, maxStep, parallel, benchmarking, groupDists ? groupDists + l * groupStride : nullptr);
        });
        tBegin += since(t0);
        for (;;) {
            t0 = now();
            bool any = false;
            for (int64_t l = 0; l < act; l++) {
                more[(size_t)l] = lanes[(size_t)l]->wsState.NextBatch(&batch[(size_t)l]) ? 1 : 0;
                any = any || more[(size_t)l];
            }
            if (!any) break;   // all lanes use the same maxStep, so they finish together
            tNext += since(t0);
            t0 = now();
            nSteps++;
            if (groupable) {
                std::vector<PIRGraphInfo *> gi;
                std::vector<const std::vector<int64_t> *> gb;
                std::vector<const float *> gq;
                std::vector<std::vector<const uint64_t *> *> ge;
                std::vector<std::vector<float> *> gd;
                for (int64_t l = 0; l < act; l++) {
                    if (!more[(size_t)l]) continue;
                    gi.push_back(infos[(size_t)l]); gb.push_back(&batch[(size_t)l]); gq.push_back(benchmarking ? nullptr : qptr[(size_t)l]);
                    ge.push_back(&rawEntries[(size_t)l]); gd.push_back(&srcDists[(size_t)l]);
                }
                if (PIRGraphInfo::FetchGroupRaw(gi, gb, gq, ge, gd) != 0) return -1;
            } else {
                for (int64_t l = 0; l < act; l++)
                    if (more[(size_t)l] && lanes[(size_t)l]->Graph->GetVertexInfoWithDist(batch[(size_t)l], benchmarking ? nullptr : qptr[(size_t)l],
                                                                                          &results[(size_t)l], &srcDists[(size_t)l]) != 0)
                        return -1;
            }
            tFetch += since(t0);
            t0 = now();
            pool.ParallelFor((size_t)act, [&](size_t l) {
                if (!more[l]) return;
                if (groupable) lanes[l]->wsState.CollectFreshRaw(batch[l], rawEntries[l], srcDists[l]);
                else lanes[l]->wsState.CollectFresh(results[l], srcDists[l]);
            });
            tCollect += since(t0);
            t0 = now();
            // distances the fetch did not provide (entries served from a lane's local cache): one launch for all lanes
            size_t nMissing = 0;
            for (int64_t l = 0; l < act; l++) {
                missBase[(size_t)l] = nMissing;
                if (more[(size_t)l]) nMissing += lanes[(size_t)l]->wsState.MissingVectors().size();
            }
            missDist.assign(nMissing, 0.f);
            if (nMissing) {
                missA.resize(nMissing * (size_t)dim);
                missB.resize(nMissing * (size_t)dim);
                for (int64_t l = 0; l < act; l++) {
                    if (!more[(size_t)l]) continue;
                    const auto &mv = lanes[(size_t)l]->wsState.MissingVectors();
                    for (size_t j = 0; j < mv.size(); j++) {
                        memcpy(&missA[(missBase[(size_t)l] + j) * (size_t)dim], mv[j], (size_t)dim * 4);
                        memcpy(&missB[(missBase[(size_t)l] + j) * (size_t)dim], qptr[(size_t)l], (size_t)dim * 4);
                    }
                }
                check(pm_l2_pairs(missA.data(), missB.data(), nMissing, (uint64_t)dim, missDist.data(), lanes[0]->Graph->Device()), "pm_l2_pairs");
            }
            tMiss += since(t0);
            nMissTotal += nMissing;
            t0 = now();
            pool.ParallelFor((size_t)act, [&](size_t l) {
                if (!more[l]) return;
                if (groupable) lanes[l]->wsState.ApplyFreshRaw(batch[l], rawEntries[l], missDist.data() + missBase[l]);
                else lanes[l]->wsState.ApplyFresh(results[l], missDist.data() + missBase[l]);
            });
            tApply += since(t0);
        }
        for (int64_t l = 0; l < act; l++)
            lanes[(size_t)l]->wsState.Finish(&(*ret)[(size_t)((base + l) * k)], &(*stepRet)[(size_t)((base + l) * k)]);
    }
    if (prof && nSteps)
        fprintf(stderr, "[lockstep profile] %lld lanes, %llu steps, us per step: next %.1f | fetch %.1f | collect %.1f | missing-dist %.1f (%.1f per step) | "
                        "apply %.1f ; begin %.1f us per round\n", (long long)L, (unsigned long long)nSteps, tNext / nSteps * 1e6, tFetch / nSteps * 1e6,
                tCollect / nSteps * 1e6, tMiss / nSteps * 1e6, (double)nMissTotal / nSteps, tApply / nSteps * 1e6, tBegin / ((nq + L - 1) / L) * 1e6);
    return 0;
}

void RobustPruneBatch(pm_db *vecDb, int64_t dim, const std::vector<int64_t> &us, const std::vector<std::vector<int64_t>> &candidates,
                      int64_t m, float alpha, std::vector<std::vector<int64_t>> *out) {
    const size_t B = us.size();
    out->assign(B, {});
    // pair list: per vertex with more than m candidates, K pairs (u, c_i) then K(K-1)/2 pairs (c_i, c_j), i < j
    std::vector<int64_t> ia, ib;
    std::vector<size_t> base(B, 0);
    for (size_t b = 0; b < B; b++) {
        const auto &c = candidates[b];
        const size_t K = c.size();
        if ((int64_t)K <= m) continue;
        base[b] = ia.size();
        for (size_t i = 0; i < K; i++) { ia.push_back(us[b]); ib.push_back(c[i]); }
        for (size_t i = 0; i < K; i++)
            for (size_t j = i + 1; j < K; j++) { ia.push_back(c[i]); ib.push_back(c[j]); }
    }
    std::vector<float> d(ia.size());
    if (!ia.empty()) check(pm_l2_idpairs(vecDb, (uint64_t)dim, ia.data(), ib.data(), ia.size(), d.data()), "pm_l2_idpairs");
    struct IdWithDist { int64_t id; float dist; size_t pos; };
    for (size_t b = 0; b < B; b++) {
        const auto &c = candidates[b];
        const size_t K = c.size();
        if ((int64_t)K <= m) { (*out)[b] = c; continue; }                       // build_graph.go:170-172
        const float *du = d.data() + base[b], *dm = du + K;
        auto pair_dist = [&](size_t i, size_t j) {                              // candidate positions, any order
            if (i > j) std::swap(i, j);
            return dm[i * K - i * (i + 1) / 2 + (j - i - 1)];
        };
        std::vector<IdWithDist> dist2u(K);
        for (size_t i = 0; i < K; i++) dist2u[i] = {c[i], du[i], i};             // :175-181
        std::stable_sort(dist2u.begin(), dist2u.end(), [](const IdWithDist &x, const IdWithDist &y) { return x.dist < y.dist; });  // :183-185
        std::vector<IdWithDist> accept, discarded;
        for (size_t i = 0; i < K; i++) {                                        // :187-208
            const float dist_uv = dist2u[i].dist;
            bool ok = true;
            for (size_t j = 0; j < accept.size(); j++) {
                const float dj = accept[j].pos == dist2u[i].pos ? 0.f : pair_dist(accept[j].pos, dist2u[i].pos);
                if (dj * alpha < dist_uv) { ok = false; break; }
            }
            if (ok) {
                accept.push_back(dist2u[i]);
                if ((int64_t)accept.size() == m) break;
            } else {
                discarded.push_back(dist2u[i]);
            }
        }
        if ((int64_t)accept.size() < m)                                         // :213-226
            for (size_t i = 0; i < discarded.size() && (int64_t)accept.size() < m; i++) accept.push_back(discarded[i]);
        for (auto &a : accept) (*out)[b].push_back(a.id);
    }
}

}  // namespace graphann
