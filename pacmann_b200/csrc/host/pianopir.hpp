// Host-side mirror of the reference's `pianopir` Go package (pianopir/pir.go, batch-pir.go, util.go),
// written in C++ because no Go toolchain exists in this image.  Same type and method names, same
// argument meaning and error behaviour; the data-parallel inner loops go through the C-ABI of
// libpacmann_cuda.so (include/pacmann_cuda.h) exactly where the cgo bridge would call it:
//
//   PianoPIRClient::Preprocessing   -> pm_hintgen + pm_gather_rows     (pir.go:267-352)
//   PianoPIRServer::PrivateQuery    -> pm_answer_batch                 (pir.go:65-88)
//   SimpleBatchPianoPIR::Preprocessing -> ONE pm_hintgen over all sub-PIRs (batch-pir.go:119-155)
//   SimpleBatchPianoPIR::Query      -> ONE pm_answer_batch for every sub-query of the call (batch-pir.go:170-248)
//   GetLongKey                      -> pm_expand_key                   (util.go:167-171)
//
// What stays on the host is what stays in Go in the drop-in: parameter derivation, hint-table
// bookkeeping, the online client's first-match hint search and set expansion (pir.go:405-427, ranked
// "next" in SURVEY.md 8f), caches and statistics.  This file only includes the public C header.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/pacmann_cuda.h"
#include "flatmap.hpp"

namespace pianopir {

constexpr uint64_t DefaultProgramPoint = 0x7fffffff;  // pir.go:15
constexpr uint64_t RealQueryPerPartition = 2;         // batch-pir.go:13
constexpr uint64_t QueryPerPartition = 2;             // batch-pir.go:14
constexpr uint64_t DefaultValue = 0xdeadbeef;         // batch-pir.go:15

struct PrfKey128 { uint8_t b[16]; };
using PrfKey = PrfKey128;

// counter-based stand-in for the reference's time-seeded math/rand draws (see pacmann_b200/keys.py)
uint64_t Mix64(uint64_t seed, uint64_t ctr);
PrfKey DeriveKey(uint64_t key_seed, uint64_t epoch, uint64_t parts, uint64_t i, uint64_t key_seed_hi = 0);
uint64_t SecureRandom64();   // OS CSPRNG (getrandom)

// util.go
std::vector<uint32_t> GetLongKey(const PrfKey128 &key);                                   // :167-171 (GPU)
uint64_t PRFEvalWithLongKeyAndTag(const std::vector<uint32_t> &longKey, uint64_t tag, uint64_t x);  // :157-165 (host AES)
void GenParams(uint64_t DBSize, uint64_t *ChunkSize, uint64_t *SetSize);                  // :97-108
void EntryXor(uint64_t *a, const uint64_t *b, uint64_t entrySize);                        // pir.go:258-265

struct PianoPIRConfig {  // pir.go:18-26
    uint64_t DBEntryByteNum, DBEntrySize, DBSize, ChunkSize, SetSize, ThreadNum, FailureProbLog2;
};

// A device-resident rawDB shared by every PianoPIR over it (the Go code aliases one []uint64, pir.go:34-38).
struct DeviceDB {
    pm_db *h = nullptr;
    uint64_t n_rows = 0, entry_u64 = 0;
    int device = 0;
    bool owned = false;
    DeviceDB(const uint64_t *rawDB, uint64_t n_rows, uint64_t entry_u64, int device);
    ~DeviceDB();
    DeviceDB(const DeviceDB &) = delete;
};

struct QueryError {
    enum Code { None = 0, OutOfRange = 1, BudgetExceeded = 2, TooManyInChunk = 3, NoHitHint = 4 };
};

class PianoPIRServer {  // pir.go:28-88
public:
    PianoPIRServer(const PianoPIRConfig *config, DeviceDB *db, uint64_t row0) : config(config), db(db), row0(row0) {}
    int NonePrivateQuery(uint64_t idx, std::vector<uint64_t> *ret);
    int PrivateQuery(const std::vector<uint32_t> &offsets, std::vector<uint64_t> *ret);
    const PianoPIRConfig *config;
    DeviceDB *db;
    uint64_t row0;
};

// One prepared (not yet answered) client query: everything Query() decides before it needs the server's
// response (pir.go:354-447).  Lets SimpleBatchPianoPIR put all server answers of a call in one launch.
struct PendingQuery {
    enum Kind { Dummy, Cached, Real, Failed } kind = Failed;
    int err = 0;
    uint64_t idx = 0, chunkId = 0, hitId = 0, inGroupIdx = 0;
    uint64_t backupTag = 0;
    std::vector<uint32_t> offsets;  // what goes to the server (Dummy, Real)
};

// Threads a lock-step group uses for its per-lane host work: PM_HOST_THREADS, else OMP_NUM_THREADS (torchrun sets it
// to 1 per rank), else min(8, cores).  Several groups (and several ranks) run side by side, so a group must not grab the
// whole machine: oversubscribed spin-waiting teams cost two orders of magnitude (measured: 8 ranks x 4 groups x 8
// threads on 32 cores = 230 ms per step instead of 1 ms).
int HostThreads();

// The pool behind those threads: HostThreads() - 1 workers owned by the calling (driver) thread plus the caller itself.
// Workers spin briefly for the next loop, then sleep.  Not OpenMP on purpose: a second OpenMP runtime in the process
// (torch brings its own) changes libgomp's spin / sleep policy under our feet (measured: 2 800 vs 4 600 queries/s).
class WorkerPool {
public:
    explicit WorkerPool(int threads);
    ~WorkerPool();
    // fn(i) for every i in [0, n), on the workers and the caller; returns when all are done; rethrows the first exception
    void ParallelFor(size_t n, const std::function<void(size_t)> &fn);
    static WorkerPool &Local();   // the calling thread's pool (created on first use)

private:
    struct Impl;
    Impl *impl;
};

// localCache (pir.go:127, :381-383, :468): idx -> entry.  Entries live in fixed-size slabs (stable addresses, reused after
// clear()) instead of one heap vector per entry: a search step inserts ~100 entries per client and with many clients in
// lock step the per-entry allocations (first-touch page faults under the process-wide mm lock) were the largest host
// cost of a step.  Reserve() allocates and touches the slabs up front.
class EntryCache {
public:
    void Init(uint64_t entryWords) { E = entryWords; clear(); }
    void Reserve(uint64_t entries);
    void clear() { slot.reset(reserved); used = 0; }
    bool has(uint64_t idx) const { return slot.has(idx); }
    const uint64_t *find(uint64_t idx) const {
        const uint64_t *s = slot.find(idx);
        return s ? at(*s) : nullptr;
    }
    const uint64_t *put(uint64_t idx, const uint64_t *entry);
    size_t size() const { return slot.size(); }

private:
    static constexpr uint64_t kPerSlab = 512;
    const uint64_t *at(uint64_t s) const { return slabs[s / kPerSlab].get() + (s % kPerSlab) * E; }
    uint64_t E = 0, used = 0, reserved = 0;
    std::vector<std::unique_ptr<uint64_t[]>> slabs;
    FlatMap slot;
};

class PianoPIRClient {  // pir.go:91-471
public:
    explicit PianoPIRClient(const PianoPIRConfig *config);
    double LocalStorageSize() const;                                                    // :178-190
    void Initialization();                                                              // :203-255
    void Preprocessing(PianoPIRServer *server);                                         // :267-301 (+UpdatePreprocessing)
    int Query(uint64_t idx, PianoPIRServer *server, bool realQuery, std::vector<uint64_t> *ret);  // :354-471
    // two-phase form of Query used for batching; PrepareQuery + FinishQuery == Query
    void PrepareQuery(uint64_t idx, bool realQuery, PendingQuery *pq);
    void FinishQuery(const PendingQuery &pq, const uint64_t *response, std::vector<uint64_t> *ret);
    // describe this client's hint table as a pm_hint_job (Initialization numbering) and install results
    void FillHintJob(uint64_t row0, pm_hint_job *job, uint64_t *parity_out) const;
    void DrawReplacementIdx(std::vector<uint64_t> *local_idx);

    const PianoPIRConfig *config;
    bool skipPrep = false;
    PrfKey masterKey{};
    std::vector<uint32_t> longKey;
    uint64_t MaxQueryNum = 0, FinishedQueryNum = 0, maxQueryPerChunk = 0;
    std::vector<uint64_t> QueryHistogram;
    uint64_t primaryHintNum = 0;
    std::vector<uint64_t> primaryShortTag, primaryParity, primaryProgramPoint;
    // [SetSize][maxQueryPerChunk(*E)] flattened (the Go code uses slices of slices, pir.go:113-118)
    std::vector<uint64_t> replacementIdx, replacementVal, backupShortTag, backupParity;
    EntryCache localCache;
    // Client secrets.  Drawn from the OS CSPRNG in the constructor (128 key-seed bits); SetSeeds() injects deterministic
    // values for the parity tests.  keyEpoch counts this client's preprocessings: every Initialization() derives a
    // fresh key from (seed, epoch, index) and advances it; prepEpoch = the epoch of the current hint table.
    void SetSeeds(uint64_t key_seed, uint64_t repl_seed);
    uint64_t keySeed = 0, keySeedHi = 0, keyEpoch = 0, prepEpoch = 0, keyIndex = 0, keyParts = 1, replSeed = 0, dummySeed = 0, dummyCtr = 0;
    std::vector<uint64_t> pendingCached;  // idx prepared in the current batch whose value arrives at Finish
};

class PianoPIR {  // pir.go:473-548
public:
    PianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, DeviceDB *db, uint64_t row0, uint64_t FailureProbLog2);
    void Preprocessing();
    void DummyPreprocessing();
    int Query(uint64_t idx, bool realQuery, std::vector<uint64_t> *ret);
    double LocalStorageSize() const { return client.LocalStorageSize(); }
    double CommCostPerQuery() const { return double(config.SetSize * 4 + config.DBEntrySize * 8); }
    const PianoPIRConfig *Config() const { return &config; }
    PianoPIRConfig config;
    PianoPIRClient client;
    PianoPIRServer server;
};

struct SimpleBatchPianoPIRConfig {  // batch-pir.go:19-28
    uint64_t DBEntryByteNum, DBEntrySize, DBSize, BatchSize, PartitionNum, PartitionSize, ThreadNum, FailureProbLog2;
};

class SimpleBatchPianoPIR {  // batch-pir.go:40-276
public:
    // rawDB is uploaded once (replicated on `device`); len_rawDB must equal DBSize*DBEntryByteNum/8 (batch-pir.go:57-59)
    SimpleBatchPianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, const uint64_t *rawDB,
                        uint64_t len_rawDB, uint64_t FailureProbLog2, int device = 0);
    // a further client over a rawDB that is already resident (one DB replica per GPU, one client per user)
    SimpleBatchPianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, DeviceDB *sharedDB, uint64_t FailureProbLog2);
    ~SimpleBatchPianoPIR();
    void SetSeeds(uint64_t keySeed, uint64_t replSeed);
    void Preprocessing();
    void DummyPreprocessing();
    int Query(const std::vector<uint64_t> &idx, std::vector<std::vector<uint64_t>> *ret);
    void RecordStats(double prepTime);
    double LocalStorageSize() const;
    uint64_t CommCostPerBatchOnline() const;
    uint64_t CommCostPerBatchOffline() const { return commCostPerBatchOffline; }
    double PreprocessingTime() const { return preprocessingTime; }
    const SimpleBatchPianoPIRConfig *Config() const { return &config; }
    std::string PrintInfo() const;

    SimpleBatchPianoPIRConfig config;
    std::vector<PianoPIR *> subPIR;
    DeviceDB *db = nullptr;
    bool ownsDB = true;
    uint64_t FinishedBatchNum = 0, QueriesMadeInPartition = 0, SupportBatchNum = 0;
    uint64_t localStorage = 0, commCostPerBatchOnline = 0, commCostPerBatchOffline = 0;
    double preprocessingTime = 0, preprocessingTotal = 0;   // last / sum over all Preprocessing() calls (the "maintenance" of private-search.go:219-240)
    uint64_t preprocessingCount = 0;
    uint64_t serverQueries = 0, serverLaunches = 0;  // accounting: sub-queries answered / pm_answer_batch calls

    // GPU-resident client (pm_client_*): hint tables stay in HBM, hint search / refresh run on the GPU.
    // Must be enabled before Preprocessing(); the host-side tables of the sub-PIRs are then only refreshed on
    // request (SyncTablesFromDevice), counters and the local cache stay on the host.
    // lanes > 1: the pm_client is created with lanes * PartitionNum parts; lane 0 is this object, further independent
    // client instances (own keys, hint tables, caches, counters) attach to it as lanes 1.. and can then be driven in
    // lock step by QueryFlatGroup -- one pm_client_query_batch_l2m call for all of them.
    void EnableResidentClient(uint32_t lanes = 1);
    void AttachResidentClient(SimpleBatchPianoPIR *owner, uint32_t lane);
    void SyncTablesFromDevice(uint64_t i);
    int QueryFlat(const uint64_t *idx, size_t n, uint64_t *out, const float *query_vec, uint64_t dim, float *dists);
    struct GroupCall {
        SimpleBatchPianoPIR *pir;
        const uint64_t *idx; size_t n; uint64_t *out;   // as QueryFlat
        const float *query_vec; float *dists;           // query_vec may be null (then no distances for this lane)
        int rc;
        // out == nullptr: no copy of the answers; out_ptrs[i] points at the entry of idx[i] instead (into the group's
        // result buffer, the lane's local cache or a zero row), valid until the lane's next Query call
        const uint64_t **out_ptrs = nullptr;
    };
    // QueryFlat of several lanes of one pm_client at once.  Every lane ends in exactly the state its own QueryFlat
    // would have left (a lane that may exhaust a sub-PIR's budget inside this call is simply run on its own).
    static int QueryFlatGroup(std::vector<GroupCall> &calls, uint64_t dim);
    // ---- device-resident search (pm_search_*): the owner of a client group also owns the search
    // object; every lane keeps the host-side counters that need no entry data (batch budget, statistics).
    pm_search *devSearch = nullptr;            // owner only
    uint64_t devSearchKey = 0;                 // (max_step, parallel, start-vertex stamp) the object was built for
    bool hostMode = false;                     // this lane left the device path (a sub-PIR budget about to run out); its local
                                               // cache lives on the host until the next Preprocessing() of the whole batch
    bool DeviceFetchAccounting(size_t n);      // batch-pir.go:239-245 for one device-built Query call; true = Preprocessing() due
    void AbsorbDeviceRound(const uint64_t *finished, uint64_t serverQ);
    bool DeviceRoundIsSafe(uint64_t maxStep, size_t n) const;   // no sub-PIR can reach its query budget within maxStep calls
    void EnterHostMode(pm_search *s);          // pull this lane's device caches into the host-side localCache
    bool resident = false;
    bool ownsClient = true;
    uint32_t partBase = 0, clientLanes = 1;
    double profQueryTotal = 0, profGpuCall = 0;  // PM_HOST_PROFILE=1 prints them when the object is destroyed
    uint64_t profQueryCalls = 0;
    pm_client *rclient = nullptr;

private:
    void Init(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, DeviceDB *theDB, bool owns, uint64_t FailureProbLog2);
    struct PendRec { uint64_t part, global, local; int kind; /* 0 dummy, 1 real, 2 cached */ int64_t qpos; };
    std::vector<std::vector<uint64_t>> wsLists;   // per-call scratch, kept to avoid reallocation
    std::vector<PendRec> wsPend;
    std::vector<pm_client_query> wsQueries;
    std::vector<uint64_t> wsOut, wsZero, wsSolo, wsCached;
    size_t wsCachedUsed = 0;
    std::vector<int32_t> wsStatus;
    std::vector<float> wsDist;
    struct Resp { const uint64_t *entry; float dist; };
    FlatMap wsResponses;           // global index -> position in wsRespList
    std::vector<Resp> wsRespList;
    std::vector<uint64_t> wsPendingReal;
    // the three pieces of QueryFlat, shared with QueryFlatGroup
    void beginCall(const uint64_t *idx, size_t n, bool *bad);
    void pushRecord(uint64_t part, uint64_t globalIdx);
    void settle(size_t pbase, const uint64_t *res, const int32_t *status, const float *dist);
    // true: the batch budget is used up, Preprocessing() is due
    bool finishCall(const uint64_t *idx, size_t n, uint64_t *out, const uint64_t **out_ptrs, float *dists);
    uint64_t *groupBuf = nullptr;   // page-locked result buffer of QueryFlatGroup calls hosted by this object (pm_host_alloc)
    size_t groupBufWords = 0;
    bool mayFlushInside(size_t n) const;
    void reserveCaches();
    void PreprocessResident(const std::vector<uint32_t> &ids, bool skipPrep);
    int QueryResident(const std::vector<uint64_t> &idx, std::vector<std::vector<uint64_t>> *ret);
    void Flush(std::vector<PendingQuery> &pend, std::vector<uint64_t> &pend_part, std::vector<uint64_t> &pend_global,
               std::unordered_map<uint64_t, std::vector<uint64_t>> &responses);
};

}  // namespace pianopir
