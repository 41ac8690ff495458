// Host-side mirror of the reference's `graphann` package (graphann/search.go, build_graph.go:119-134)
// and of the PIRGraphInfo adapter in private-search.go:334-531, in C++ (no Go toolchain here).
// Distance evaluation goes through the C-ABI (pm_l2_query / pm_l2_batch); traversal control stays on
// the host as it stays in Go.  Tie rules where Go leaves the order unspecified:
// start-vertex ranking is stable in input order, the explore queue is container/heap's binary heap,
// the final ranking orders by (distance, id).
#pragma once
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "pianopir.hpp"

namespace graphann {

struct Vertex {  // search.go:12-16
    int64_t Id = 0;
    std::vector<int64_t> Neighbors;
    std::vector<float> Vector;
};

class GetGraphInfo {  // search.go:20-25
public:
    virtual ~GetGraphInfo() {}
    virtual void Preprocess() = 0;
    virtual void GetMetadata(int64_t *n, int64_t *dim, int64_t *m) const = 0;
    virtual int GetVertexInfo(const std::vector<int64_t> &ids, std::vector<Vertex> *out) = 0;
    virtual int GetStartVertex(std::vector<Vertex> *out) = 0;
    // GetVertexInfo plus, where the source can provide it for free, dists[i] = L2Dist(out[i].Vector, query);
    // NaN = not provided (the caller evaluates L2Dist itself).  Default: no distances.
    virtual int GetVertexInfoWithDist(const std::vector<int64_t> &ids, const float *query, std::vector<Vertex> *out,
                                      std::vector<float> *dists);
    virtual int Device() const { return 0; }
};

// L2Dist (build_graph.go:119-127) for one pair: a single-distance launch, kept for API parity
float L2Dist(const std::vector<float> &v1, const std::vector<float> &v2, int device = 0);

class BasicGraphInfo : public GetGraphInfo {  // search.go:29-65
public:
    BasicGraphInfo(int64_t N, int64_t Dim, int64_t M, const int32_t *graph, const float *vectors)
        : N(N), Dim(Dim), M(M), Graph(graph), Vectors(vectors) {}
    void Preprocess() override {}
    void GetMetadata(int64_t *n, int64_t *dim, int64_t *m) const override { *n = N; *dim = Dim; *m = M; }
    int GetVertexInfo(const std::vector<int64_t> &ids, std::vector<Vertex> *out) override;
    int GetStartVertex(std::vector<Vertex> *out) override;
    int64_t N, Dim, M;
    const int32_t *Graph;   // [N][M]
    const float *Vectors;   // [N][Dim]
};

// PIRGraphInfo (private-search.go:334-531): the graph behind a SimpleBatchPianoPIR.
class PIRGraphInfo : public GetGraphInfo {
public:
    PIRGraphInfo(int64_t N, int64_t Dim, int64_t M, const int32_t *graph, const float *vectors, bool skipPrep,
                 bool nonPrivate, uint64_t seed, int device = 0);
    ~PIRGraphInfo() override;
    void Preprocess() override;                                                        // :355-412
    void GetMetadata(int64_t *n, int64_t *dim, int64_t *m) const override { *n = N; *dim = Dim; *m = M; }
    int GetVertexInfo(const std::vector<int64_t> &ids, std::vector<Vertex> *out) override;  // :441-506
    int GetVertexInfoWithDist(const std::vector<int64_t> &ids, const float *query, std::vector<Vertex> *out,
                              std::vector<float> *dists) override;
    int GetStartVertex(std::vector<Vertex> *out) override;                             // :508-531
    int Device() const override { return device; }
    int64_t N, Dim, M;
    const int32_t *graph;
    const float *vectors;
    bool skipPrep, NonPrivateMode;
    bool residentClient = true;  // keep the hint tables on the GPU (pm_client_*)
    PIRGraphInfo *shareDBWith = nullptr;  // another client's graph info whose resident rawDB this one reuses (one per user)
    // Client groups for lock-step search: the group's first client sets groupLanes = L before Preprocess(); clients
    // 1..L-1 set laneOf = the first client and lane = their number (they also share its rawDB).  All lanes then live in
    // one pm_client and FetchGroupRaw fetches for all of them with one device call.
    uint32_t groupLanes = 1, lane = 0;
    PIRGraphInfo *laneOf = nullptr;
    static int FetchGroupRaw(const std::vector<PIRGraphInfo *> &infos, const std::vector<const std::vector<int64_t> *> &ids,
                             const std::vector<const float *> &queries, const std::vector<std::vector<const uint64_t *> *> &entries,
                             const std::vector<std::vector<float> *> &dists);
    void unpackResponses(const std::vector<int64_t> &ids, std::vector<Vertex> *out);
    uint64_t DBEntryByteNum = 0, DBTotalSize = 0;
    std::vector<uint64_t> rawDB;
    pianopir::SimpleBatchPianoPIR *PIR = nullptr;
    int64_t totalQueryNum = 0, succQueryNum = 0;
    std::vector<uint64_t> wsIdx, wsResp;  // per-call scratch
    uint64_t seed;
    int device;
};

// Entry2VectorAndNeighbors (private-search.go:418-439) and its inverse (the packing loop :371-397)
void Entry2VectorAndNeighbors(int64_t dim, int64_t m, const uint64_t *entry, std::vector<float> *vector,
                              std::vector<int64_t> *neighbors);
void PackEntry(int64_t dim, int64_t m, const float *vector, const int32_t *neighbors, uint64_t *entry);

struct VD { float dist; int64_t id; };
struct ExploreQueue {   // container/heap over (dist, id), search.go:92-111
    std::vector<VD> a;
    void up(int64_t j);
    void down(int64_t i0, int64_t n);
    void Push(VD v);
    VD Pop();
    size_t Len() const { return a.size(); }
};

class GraphANNFrontend;
// One SearchKNN in progress (search.go:114-234): Begin, then NextBatch -> fetch -> Consume once per step, then Finish.
// SearchKNN drives it alone, SearchKNNLockstep drives one per lane with a shared fetch.
class SearchState {
public:
    // startDists: distances of the frontend's start vertices to the query if the caller already has them (lock-step
    // driver: one launch for all lanes), else nullptr
    void Begin(GraphANNFrontend *front, const float *queryVector, int64_t k, int64_t maxStep, int64_t parallel, bool benchmarking,
               const float *startDists = nullptr);
    bool NextBatch(std::vector<int64_t> *batchQ);
    void Consume(const std::vector<Vertex> &queryResults, const std::vector<float> &srcDists);
    // Consume in parts: CollectFresh, then L2Dist(MissingVectors()[j], query) for every j, then ApplyFresh
    void CollectFresh(const std::vector<Vertex> &queryResults, const std::vector<float> &srcDists);
    const std::vector<const float *> &MissingVectors() const { return ptrs; }
    void ApplyFresh(const std::vector<Vertex> &queryResults, const float *missingDists);
    // the same over entries in the wire format of private-search.go:355-439, read in place
    void CollectFreshRaw(const std::vector<int64_t> &ids, const std::vector<const uint64_t *> &entries, const std::vector<float> &srcDists);
    void ApplyFreshRaw(const std::vector<int64_t> &ids, const std::vector<const uint64_t *> &entries, const float *missingDists);
    void Finish(int64_t *ret, int64_t *stepRet);   // [k] each, -1 padded

private:
    void addKnown(const Vertex &v, float dist, int64_t step);
    template <class A> void collect(const A &res, const std::vector<float> &srcDists);
    template <class A> void apply(const A &res, const float *missingDists);
    GraphANNFrontend *f = nullptr;
    const float *queryVector = nullptr;
    int64_t k = 0, maxStep = 0, parallel = 0, n = 0, dim = 0, m = 0, step = 0;
    bool benchmarking = false;
    int device = 0;
    pianopir::FlatMap slotOf;   // vertex id -> slot in knownId / knownDist / knownStep / nbrPool
    std::vector<int64_t> knownId, knownStep, nbrPool;
    std::vector<float> knownDist, dists;
    ExploreQueue toBeExplored;
    uint64_t rseed = 0, rctr = 0;
    std::vector<size_t> fresh, missing, startOrder;
    std::vector<const float *> ptrs;
};

class GraphANNFrontend {  // search.go:69-245
public:
    explicit GraphANNFrontend(GetGraphInfo *g) : Graph(g) {}
    ~GraphANNFrontend();
    void Preprocess();
    void UploadStartVertices();   // call again after editing StartVertices by hand
    void GetMetadata(int64_t *n, int64_t *dim, int64_t *m) const { Graph->GetMetadata(n, dim, m); }
    // returns the k nearest ids and the step at which each was reached (-1 padding), search.go:114-234
    int SearchKNN(const float *queryVector, int64_t k, int64_t maxStep, int64_t parallel, bool benchmarking,
                  std::vector<int64_t> *ret, std::vector<int64_t> *stepRet);
    int SearchKNNBatch(const float *queryVectors, int64_t nq, int64_t k, int64_t maxStep, int64_t parallel,
                       bool benchmarking, std::vector<int64_t> *ret, std::vector<int64_t> *stepRet);
    GetGraphInfo *Graph;
    std::vector<Vertex> StartVertices;
    uint64_t randSeed = 0;   // stands in for Go's global math/rand in the "random query" branch (search.go:155-159)
    uint64_t queryCounter = 0;

    void StartDistances(const float *queryVector, int64_t dim, int device, std::vector<float> *out);
    SearchState wsState;                 // the search in progress (one at a time per frontend)
    const float *GroupStartDistances(const std::vector<GraphANNFrontend *> &lanes, const float *queries, int64_t act, int64_t dim, size_t *stride);
    uint64_t startVersion = 0;           // bumped whenever the start vertices are (re)uploaded

private:
    std::vector<int64_t> wsBatch;
    pm_db *startDb = nullptr;            // start vertices' vectors, resident on the GPU
    pm_db *groupStartDb = nullptr;       // first frontend of a lock-step group: the start vertices of all lanes
    uint64_t groupStamp = 0;
    std::vector<int64_t> groupIds;
    std::vector<float> groupDists;
    std::vector<int64_t> startIds;       // 0..n_start-1
    std::vector<Vertex> wsResults;       // per-step scratch reused across steps and searches
    std::vector<float> wsSrcDists;
};

// robustPrune (build_graph.go:169-236) for a batch of vertices over vectors resident on the GPU (`vecDb`: rows of dim
// fp32 viewed as dim/2 uint64).  All distances the reference evaluates one L2Dist call at a time -- u to each candidate
// and candidate to candidate -- come from ONE pm_l2_idpairs launch for the whole batch; the greedy selection then runs
// on the host over that matrix, with the same comparisons in the same order.  Candidates at equal distance from u keep
// candidate order (Go's sort.Slice leaves that order unspecified).  out[b] = pruned neighbour list of us[b].
void RobustPruneBatch(pm_db *vecDb, int64_t dim, const std::vector<int64_t> &us, const std::vector<std::vector<int64_t>> &candidates,
                      int64_t m, float alpha, std::vector<std::vector<int64_t>> *out);

// process-wide counters of the device-resident search path: rounds, queries searched on the device, queries of lanes that
// had left the device path (host mode)
void DeviceSearchStats(uint64_t out[3]);

// Lock-step SearchKNNBatch over several frontends ("lanes"): query i is searched by lane i % L; results are those of
// each lane's own SearchKNNBatch over its queries, the per-step fetches of all lanes share one device call.
int SearchKNNLockstep(const std::vector<GraphANNFrontend *> &lanes, const float *queryVectors, int64_t nq, int64_t k, int64_t maxStep,
                      int64_t parallel, bool benchmarking, std::vector<int64_t> *ret, std::vector<int64_t> *stepRet);

}  // namespace graphann
