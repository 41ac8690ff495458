// Host-side mirror of the reference's `pianopir` package over the C-ABI.  See pianopir.hpp.
#include "pianopir.hpp"

#include <immintrin.h>
#include <sys/random.h>

#include <cerrno>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace pianopir {

static void check(int rc, const char *what) {
    if (rc != PM_OK) {
        // the reference log.Fatalf's on unrecoverable conditions; here: throw, never fall back to the CPU
        throw std::runtime_error(std::string(what) + ": " + pm_last_error());
    }
}

// ---------------------------------------------------------------------------------------------
// deterministic randomness
// ---------------------------------------------------------------------------------------------
uint64_t Mix64(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// 64 bits from the OS CSPRNG (getrandom(2)): the default source of every client secret
uint64_t SecureRandom64() {
    uint64_t v = 0;
    size_t got = 0;
    while (got < sizeof(v)) {
        ssize_t r = getrandom((char *)&v + got, sizeof(v) - got, 0);
        if (r < 0) {
            if (errno == EINTR) continue;
            throw std::runtime_error("getrandom failed: no secure randomness for the client keys");
        }
        got += (size_t)r;
    }
    return v;
}
// PrfKey128 of sub-PIR i at its `epoch`-th preprocessing (the analogue of RandKey128's two rng.Uint64() draws,
// util.go:25-31).  key_seed_hi = 0 with injected (test) seeds; otherwise the second 64 secret bits, so that a default
// client's AES key carries 128 bits of OS entropy.
PrfKey DeriveKey(uint64_t key_seed, uint64_t epoch, uint64_t parts, uint64_t i, uint64_t key_seed_hi) {
    PrfKey k;
    uint64_t a = Mix64(key_seed, 2 * (epoch * parts + i)), b = Mix64(key_seed ^ key_seed_hi, 2 * (epoch * parts + i) + 1);
    memcpy(k.b, &a, 8);
    memcpy(k.b + 8, &b, 8);
    return k;
}

// ---------------------------------------------------------------------------------------------
// host AES-MMO for the ONLINE client path (first-match hint search / set expansion, pir.go:405-427),
// which stays on the host in the drop-in exactly as it stays in Go + aes_amd64.s.  The offline
// hint generation never comes here: it runs in pm_hintgen on the GPU.
// ---------------------------------------------------------------------------------------------
static uint8_t g_sbox[256];
static bool g_sbox_ready = false;
static uint8_t gmul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) p ^= a;
        bool hi = a & 0x80;
        a = (uint8_t)(a << 1);
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}
static void init_sbox() {
    if (g_sbox_ready) return;
    uint8_t p = 1, q = 1;  // generator walk: p runs over the field, q is its inverse
    do {
        p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));
        q ^= (uint8_t)(q << 1); q ^= (uint8_t)(q << 2); q ^= (uint8_t)(q << 4);
        if (q & 0x80) q ^= 0x09;
        uint8_t x = (uint8_t)(q ^ ((q << 1) | (q >> 7)) ^ ((q << 2) | (q >> 6)) ^ ((q << 3) | (q >> 5)) ^ ((q << 4) | (q >> 4)));
        g_sbox[p] = x ^ 0x63;
    } while (p != 1);
    g_sbox[0] = 0x63;
    g_sbox_ready = true;
}
static void mmo_portable(const uint32_t *rk, uint8_t in[16], uint8_t out[16]) {
    init_sbox();
    const uint8_t *k = (const uint8_t *)rk;
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ k[i];
    for (int r = 1; r <= 10; r++) {
        for (int c = 0; c < 4; c++)
            for (int row = 0; row < 4; row++) t[4 * c + row] = g_sbox[s[4 * ((c + row) & 3) + row]];
        if (r < 10) {
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = gmul(a0, 2) ^ gmul(a1, 3) ^ a2 ^ a3;
                s[4 * c + 1] = a0 ^ gmul(a1, 2) ^ gmul(a2, 3) ^ a3;
                s[4 * c + 2] = a0 ^ a1 ^ gmul(a2, 2) ^ gmul(a3, 3);
                s[4 * c + 3] = gmul(a0, 3) ^ a1 ^ a2 ^ gmul(a3, 2);
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; i++) s[i] ^= k[16 * r + i];
    }
    for (int i = 0; i < 16; i++) out[i] = s[i] ^ in[i];
}
__attribute__((target("aes,sse4.1"))) static inline uint64_t prf_aesni(const uint32_t *rk, uint64_t tag, uint64_t x) {
    const __m128i *k = (const __m128i *)rk;
    __m128i in = _mm_set_epi64x(0, (long long)((tag << 35) + x));
    __m128i s = _mm_xor_si128(in, _mm_loadu_si128(k));
    for (int r = 1; r < 10; r++) s = _mm_aesenc_si128(s, _mm_loadu_si128(k + r));
    s = _mm_aesenclast_si128(s, _mm_loadu_si128(k + 10));
    return (uint64_t)_mm_cvtsi128_si64(_mm_xor_si128(s, in));
}
static bool has_aesni() {
    static int v = -1;
    if (v < 0) {
        __builtin_cpu_init();
        v = __builtin_cpu_supports("aes") && __builtin_cpu_supports("sse4.1");
    }
    return v;
}
static inline uint64_t prf_host(const uint32_t *rk, uint64_t tag, uint64_t x) {
    if (has_aesni()) return prf_aesni(rk, tag, x);
    uint8_t in[16] = {0}, out[16];
    uint64_t v = (tag << 35) + x;
    memcpy(in, &v, 8);
    mmo_portable(rk, in, out);
    uint64_t o;
    memcpy(&o, out, 8);
    return o;
}

std::vector<uint32_t> GetLongKey(const PrfKey128 &key) {
    std::vector<uint32_t> rk(44);
    check(pm_expand_key(key.b, rk.data()), "GetLongKey/pm_expand_key");
    return rk;
}
uint64_t PRFEvalWithLongKeyAndTag(const std::vector<uint32_t> &longKey, uint64_t tag, uint64_t x) {
    return prf_host(longKey.data(), tag, x);
}
void GenParams(uint64_t DBSize, uint64_t *ChunkSize, uint64_t *SetSize) {
    uint64_t target = (uint64_t)(2 * std::sqrt((double)DBSize));
    uint64_t c = 1;
    while (c < target) c *= 2;
    uint64_t s = (uint64_t)std::ceil((double)DBSize / (double)c);
    s = (s + 3) / 4 * 4;
    *ChunkSize = c;
    *SetSize = s;
}
// client-side single-entry XOR; same truncation as xorSlices (4 words per step, tail untouched)
void EntryXor(uint64_t *a, const uint64_t *b, uint64_t entrySize) {
    const uint64_t n4 = entrySize & ~3ull;
    for (uint64_t i = 0; i < n4; i++) a[i] ^= b[i];
}

// ---------------------------------------------------------------------------------------------
DeviceDB::DeviceDB(const uint64_t *rawDB, uint64_t n_rows_, uint64_t entry_u64_, int device_)
    : n_rows(n_rows_), entry_u64(entry_u64_), device(device_), owned(true) {
    check(pm_db_create(rawDB, n_rows, entry_u64, device, &h), "pm_db_create");
}
DeviceDB::~DeviceDB() {
    if (owned && h) pm_db_destroy(h);
}

// ---------------------------------------------------------------------------------------------
// server
// ---------------------------------------------------------------------------------------------
int PianoPIRServer::NonePrivateQuery(uint64_t idx, std::vector<uint64_t> *ret) {
    ret->assign(config->DBEntrySize, 0);
    if (idx >= config->DBSize) return idx < config->ChunkSize * config->SetSize ? 0 : QueryError::OutOfRange;
    check(pm_gather_rows(db->h, row0, config->DBSize, &idx, 1, ret->data()), "pm_gather_rows");
    return 0;
}
int PianoPIRServer::PrivateQuery(const std::vector<uint32_t> &offsets, std::vector<uint64_t> *ret) {
    ret->assign(config->DBEntrySize, 0);
    uint64_t r0 = row0, nr = config->DBSize;
    uint32_t c = (uint32_t)config->ChunkSize, s = (uint32_t)config->SetSize;
    check(pm_answer_batch(db->h, &r0, &nr, &c, &s, offsets.data(), config->SetSize, 1, ret->data()), "pm_answer_batch");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// client
// ---------------------------------------------------------------------------------------------
static uint64_t primaryNumParam(double /*Q*/, double ChunkSize, uint64_t target) {  // pir.go:124-127
    uint64_t k = (uint64_t)std::ceil(std::log(2.0) * (double)target);
    return k * (uint64_t)ChunkSize;
}

PianoPIRClient::PianoPIRClient(const PianoPIRConfig *cfg) : config(cfg) {  // pir.go:130-175
    // secrets come from the OS CSPRNG unless a test injects seeds (SetSeeds): the reference draws its key from a
    // time-seeded rng in every Initialization (pir.go:208-211) and its dummy offsets from the global rng (pir.go:366)
    keySeed = SecureRandom64();
    keySeedHi = SecureRandom64();
    replSeed = SecureRandom64();
    dummySeed = SecureRandom64();
    MaxQueryNum = (uint64_t)(std::sqrt((double)cfg->DBSize) * std::log((double)cfg->DBSize));
    primaryHintNum = primaryNumParam((double)MaxQueryNum, (double)cfg->ChunkSize, cfg->FailureProbLog2 + 1);
    primaryHintNum = (primaryHintNum + cfg->ThreadNum - 1) / cfg->ThreadNum * cfg->ThreadNum;
    maxQueryPerChunk = 3 * (uint64_t)((double)MaxQueryNum / (double)cfg->SetSize);
    maxQueryPerChunk = (maxQueryPerChunk + cfg->ThreadNum - 1) / cfg->ThreadNum * cfg->ThreadNum;
    QueryHistogram.assign(cfg->SetSize, 0);
}

double PianoPIRClient::LocalStorageSize() const {
    double s = 0, P = (double)primaryHintNum, EB = (double)config->DBEntryByteNum;
    s += P * 8 + P * EB + P * 8;
    double B = (double)config->SetSize * (double)maxQueryPerChunk;
    s += B * 8 + B * EB + B * 8 + B * EB;
    return s;
}

void PianoPIRClient::Initialization() {
    FinishedQueryNum = 0;
    // a fresh key for EVERY preprocessing of this client (pir.go:208-211), also the budget-triggered ones: the epoch
    // advances here, so no hint table is ever rebuilt under a key that has already been shown to the server
    prepEpoch = keyEpoch++;
    masterKey = DeriveKey(keySeed, prepEpoch, keyParts, keyIndex, keySeedHi);
    longKey = GetLongKey(masterKey);
    const uint64_t S = config->SetSize, M = maxQueryPerChunk, E = config->DBEntrySize, P = primaryHintNum;
    QueryHistogram.assign(S, 0);
    uint64_t shortTagCount = 0;
    primaryShortTag.resize(P);
    primaryParity.assign(P * E, 0);
    primaryProgramPoint.assign(P, DefaultProgramPoint);
    for (uint64_t i = 0; i < P; i++) primaryShortTag[i] = shortTagCount++;
    replacementIdx.assign(S * M, DefaultProgramPoint);
    replacementVal.assign(S * M * E, 0);
    backupShortTag.resize(S * M);
    backupParity.assign(S * M * E, 0);
    for (uint64_t i = 0; i < S * M; i++) backupShortTag[i] = shortTagCount++;
    localCache.Init(E);
    pendingCached.clear();
}

// Deterministic seeds: the explicit TEST hook (parity against the oracle needs the same keys on both sides).  The dummy
// stream is derived from the client's own secret seed and index, never from a public constant.
void PianoPIRClient::SetSeeds(uint64_t key_seed, uint64_t repl_seed) {
    const uint64_t ds = Mix64(Mix64(key_seed, 0xD00D), keyIndex);
    if (ds != dummySeed) { dummySeed = ds; dummyCtr = 0; }
    keySeed = key_seed;
    keySeedHi = 0;
    replSeed = repl_seed;
}

void PianoPIRClient::FillHintJob(uint64_t row0, pm_hint_job *job, uint64_t *parity_out) const {
    memset(job, 0, sizeof(*job));
    job->row0 = row0;
    job->n_rows = config->DBSize;
    job->chunk_size = config->ChunkSize;
    job->set_size = config->SetSize;
    memcpy(job->rk, longKey.data(), sizeof(job->rk));
    job->hint_begin = 0;
    job->n_hints = primaryHintNum + config->SetSize * maxQueryPerChunk;
    job->n_primary = primaryHintNum;
    job->backup_group = maxQueryPerChunk;
    job->tags = nullptr;        // Initialization numbering: tag == hint number (pir.go:220-251)
    job->skip_chunk = nullptr;  // backup group g skips chunk g (pir.go:332-334)
    job->parity_out = parity_out;
}

// replacement indices for every (chunk, slot): pir.go:345-347 with a counter-based draw
void PianoPIRClient::DrawReplacementIdx(std::vector<uint64_t> *local_idx) {
    const uint64_t S = config->SetSize, M = maxQueryPerChunk, C = config->ChunkSize;
    const uint64_t seed = Mix64(replSeed, prepEpoch * keyParts + keyIndex);
    local_idx->resize(S * M);
    for (uint64_t c = 0; c < S; c++)
        for (uint64_t j = 0; j < M; j++) {
            uint64_t off = Mix64(seed, c * M + j) & (C - 1);
            replacementIdx[c * M + j] = off + c * C;
            (*local_idx)[c * M + j] = off + c * C;
        }
}

void PianoPIRClient::Preprocessing(PianoPIRServer *server) {
    Initialization();
    if (skipPrep) return;
    const uint64_t E = config->DBEntrySize, P = primaryHintNum, B = config->SetSize * maxQueryPerChunk;
    // primary and backup parities are contiguous in hint-number order: one job, two destination tables
    std::vector<uint64_t> all((P + B) * E);
    pm_hint_job job;
    FillHintJob(server->row0, &job, all.data());
    check(pm_hintgen(server->db->h, &job, 1), "pm_hintgen");
    memcpy(primaryParity.data(), all.data(), P * E * 8);
    memcpy(backupParity.data(), all.data() + P * E, B * E * 8);
    std::vector<uint64_t> ridx;
    DrawReplacementIdx(&ridx);
    check(pm_gather_rows(server->db->h, server->row0, config->DBSize, ridx.data(), ridx.size(), replacementVal.data()),
          "pm_gather_rows");
}

void PianoPIRClient::PrepareQuery(uint64_t idx, bool realQuery, PendingQuery *pq) {
    const uint64_t S = config->SetSize, C = config->ChunkSize, M = maxQueryPerChunk;
    pq->idx = idx;
    pq->err = 0;
    pq->offsets.clear();
    if (!realQuery) {  // pir.go:363-371
        pq->kind = PendingQuery::Dummy;
        pq->offsets.resize(S);
        for (uint64_t i = 0; i < S; i++) pq->offsets[i] = (uint32_t)(Mix64(dummySeed, dummyCtr++) & (C - 1));
        return;
    }
    if (idx >= config->DBSize) {  // the reference log.Fatalf's here (pir.go:373-378)
        pq->kind = PendingQuery::Failed;
        pq->err = QueryError::OutOfRange;
        return;
    }
    bool cached = localCache.has(idx);
    for (size_t i = 0; !cached && i < pendingCached.size(); i++) cached = pendingCached[i] == idx;
    if (cached) {  // pir.go:381-383
        pq->kind = PendingQuery::Cached;
        return;
    }
    pq->kind = PendingQuery::Failed;
    if (FinishedQueryNum >= MaxQueryNum) { pq->err = QueryError::BudgetExceeded; return; }
    const uint64_t chunkId = idx / C, offset = idx % C;
    if (QueryHistogram[chunkId] >= M) { pq->err = QueryError::TooManyInChunk; return; }

    const uint32_t *rk = longKey.data();
    uint64_t hitId = DefaultProgramPoint;
    for (uint64_t i = 0; i < primaryHintNum; i++) {  // pir.go:405-414
        uint64_t hintOffset = prf_host(rk, primaryShortTag[i], chunkId) & (C - 1);
        if (hintOffset == offset) {
            if (primaryProgramPoint[i] == DefaultProgramPoint || (primaryProgramPoint[i] / C != chunkId)) {
                hitId = i;
                break;
            }
        }
    }
    if (hitId == DefaultProgramPoint) { pq->err = QueryError::NoHitHint; return; }

    std::vector<uint64_t> querySet(S);
    for (uint64_t i = 0; i < S; i++) querySet[i] = i * C + (prf_host(rk, primaryShortTag[hitId], i) & (C - 1));
    if (primaryProgramPoint[hitId] != DefaultProgramPoint)
        querySet[primaryProgramPoint[hitId] / C] = primaryProgramPoint[hitId];
    const uint64_t inGroupIdx = QueryHistogram[chunkId];
    querySet[chunkId] = replacementIdx[chunkId * M + inGroupIdx];
    pq->offsets.resize(S);
    for (uint64_t i = 0; i < S; i++) pq->offsets[i] = (uint32_t)(querySet[i] & (C - 1));

    pq->kind = PendingQuery::Real;
    pq->chunkId = chunkId;
    pq->hitId = hitId;
    pq->inGroupIdx = inGroupIdx;
    // the response-independent half of the refresh (pir.go:460-468): later queries of the same batch must see it
    primaryShortTag[hitId] = backupShortTag[chunkId * M + inGroupIdx];
    primaryProgramPoint[hitId] = idx;
    FinishedQueryNum += 1;
    QueryHistogram[chunkId] += 1;
    pendingCached.push_back(idx);
}

void PianoPIRClient::FinishQuery(const PendingQuery &pq, const uint64_t *response, std::vector<uint64_t> *ret) {
    const uint64_t E = config->DBEntrySize, M = maxQueryPerChunk;
    ret->assign(E, 0);
    switch (pq.kind) {
    case PendingQuery::Dummy:
    case PendingQuery::Failed:
        return;
    case PendingQuery::Cached:
        {
            const uint64_t *e = localCache.find(pq.idx);
            if (!e) throw std::runtime_error("local cache entry vanished");
            ret->assign(e, e + E);
        }
        return;
    case PendingQuery::Real:
        break;
    }
    memcpy(ret->data(), response, E * 8);
    const uint64_t slot = pq.chunkId * M + pq.inGroupIdx;
    EntryXor(ret->data(), &replacementVal[slot * E], E);             // pir.go:451
    EntryXor(ret->data(), &primaryParity[pq.hitId * E], E);          // pir.go:453
    memcpy(&primaryParity[pq.hitId * E], &backupParity[slot * E], E * 8);  // pir.go:461
    EntryXor(&primaryParity[pq.hitId * E], ret->data(), E);          // pir.go:463
    localCache.put(pq.idx, ret->data());                             // pir.go:468
    for (size_t i = 0; i < pendingCached.size(); i++)
        if (pendingCached[i] == pq.idx) { pendingCached.erase(pendingCached.begin() + (long)i); break; }
}

int PianoPIRClient::Query(uint64_t idx, PianoPIRServer *server, bool realQuery, std::vector<uint64_t> *ret) {
    PendingQuery pq;
    PrepareQuery(idx, realQuery, &pq);
    std::vector<uint64_t> response;
    if (pq.kind == PendingQuery::Dummy || pq.kind == PendingQuery::Real) server->PrivateQuery(pq.offsets, &response);
    FinishQuery(pq, response.data(), ret);
    return pq.err;
}

// ---------------------------------------------------------------------------------------------
// PianoPIR
// ---------------------------------------------------------------------------------------------
static PianoPIRConfig make_config(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t FailureProbLog2) {  // pir.go:479-504
    PianoPIRConfig c;
    c.DBEntryByteNum = DBEntryByteNum;
    c.DBEntrySize = DBEntryByteNum / 8;
    c.DBSize = DBSize;
    GenParams(DBSize, &c.ChunkSize, &c.SetSize);
    c.ThreadNum = 8;
    c.FailureProbLog2 = FailureProbLog2;
    return c;
}
PianoPIR::PianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, DeviceDB *db, uint64_t row0, uint64_t FailureProbLog2)
    : config(make_config(DBSize, DBEntryByteNum, FailureProbLog2)), client(&config), server(&config, db, row0) {
    if (row0 + DBSize > db->n_rows || db->entry_u64 != config.DBEntrySize)
        throw std::runtime_error("Piano PIR: rawDB slice does not match DBSize*DBEntrySize");  // pir.go:483-485
}
void PianoPIR::Preprocessing() { client.Preprocessing(&server); }
void PianoPIR::DummyPreprocessing() {
    client.Initialization();
    client.skipPrep = true;
}
int PianoPIR::Query(uint64_t idx, bool realQuery, std::vector<uint64_t> *ret) {
    if (client.FinishedQueryNum == client.MaxQueryNum) client.Preprocessing(&server);  // pir.go:527-530
    return client.Query(idx, &server, realQuery, ret);
}

// ---------------------------------------------------------------------------------------------
// SimpleBatchPianoPIR
// ---------------------------------------------------------------------------------------------
SimpleBatchPianoPIR::SimpleBatchPianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, const uint64_t *rawDB,
                                         uint64_t len_rawDB, uint64_t FailureProbLog2, int device) {
    if (len_rawDB != DBSize * (DBEntryByteNum / 8)) throw std::runtime_error("BatchPIR: len(rawDB) != DBSize*DBEntrySize");  // batch-pir.go:57-59
    Init(DBSize, DBEntryByteNum, BatchSize, new DeviceDB(rawDB, DBSize, DBEntryByteNum / 8, device), true, FailureProbLog2);
}
SimpleBatchPianoPIR::SimpleBatchPianoPIR(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, DeviceDB *sharedDB,
                                         uint64_t FailureProbLog2) {
    if (!sharedDB || sharedDB->n_rows != DBSize || sharedDB->entry_u64 != DBEntryByteNum / 8)
        throw std::runtime_error("BatchPIR: shared rawDB does not match DBSize*DBEntrySize");
    Init(DBSize, DBEntryByteNum, BatchSize, sharedDB, false, FailureProbLog2);
}
void SimpleBatchPianoPIR::Init(uint64_t DBSize, uint64_t DBEntryByteNum, uint64_t BatchSize, DeviceDB *theDB, bool owns,
                               uint64_t FailureProbLog2) {
    const uint64_t E = DBEntryByteNum / 8;
    const int device = theDB->device;
    (void)device;
    config.DBEntryByteNum = DBEntryByteNum;
    config.DBEntrySize = E;
    config.DBSize = DBSize;
    config.BatchSize = BatchSize;
    config.PartitionNum = BatchSize / RealQueryPerPartition;
    config.PartitionSize = (DBSize + config.PartitionNum - 1) / config.PartitionNum;
    config.ThreadNum = 1;
    config.FailureProbLog2 = FailureProbLog2;
    db = theDB;
    ownsDB = owns;
    for (uint64_t i = 0; i < config.PartitionNum; i++) {
        uint64_t start = i * config.PartitionSize, end = std::min((i + 1) * config.PartitionSize, DBSize);
        PianoPIR *p = new PianoPIR(end - start, DBEntryByteNum, db, start, FailureProbLog2);
        p->client.keyParts = config.PartitionNum;
        p->client.keyIndex = i;
        subPIR.push_back(p);
    }
}
SimpleBatchPianoPIR::~SimpleBatchPianoPIR() {
    if (getenv("PM_HOST_PROFILE") && profQueryCalls)
        fprintf(stderr, "[host profile] Query calls %llu: %.1f us per call, of which pm_client_query_batch %.1f us\n",
                (unsigned long long)profQueryCalls, profQueryTotal / profQueryCalls * 1e6, profGpuCall / profQueryCalls * 1e6);
    if (devSearch) pm_search_destroy(devSearch);
    if (rclient && ownsClient) pm_client_destroy(rclient);
    if (groupBuf) pm_host_free(groupBuf);
    for (auto *p : subPIR) delete p;
    if (ownsDB) delete db;
}
void SimpleBatchPianoPIR::SetSeeds(uint64_t keySeed, uint64_t replSeed) {
    for (auto *p : subPIR) p->client.SetSeeds(keySeed, replSeed);
}

std::string SimpleBatchPianoPIR::PrintInfo() const {  // batch-pir.go:95-108
    char buf[1024];
    double DBSizeInBytes = (double)config.DBSize * (double)config.DBEntryByteNum;
    uint64_t maxQuery = subPIR[0]->client.MaxQueryNum / QueryPerPartition;
    snprintf(buf, sizeof(buf),
             "-----------BatchPIR config --------\nDB size in MB = %g\nDBSize: %llu, DBEntryByteNum: %llu, BatchSize: %llu, "
             "PartitionNum: %llu, PartitionSize: %llu, ThreadNum: %llu, FailureProbLog2: %llu\nmax query num = %llu\n"
             "max query per chunk = %llu\ntotal storage = %g MB\ncomm cost per batch = %llu KB\n-----------------------------\n",
             DBSizeInBytes / 1024 / 1024, (unsigned long long)config.DBSize, (unsigned long long)config.DBEntryByteNum,
             (unsigned long long)config.BatchSize, (unsigned long long)config.PartitionNum,
             (unsigned long long)config.PartitionSize, (unsigned long long)config.ThreadNum,
             (unsigned long long)config.FailureProbLog2, (unsigned long long)maxQuery,
             (unsigned long long)subPIR[0]->client.maxQueryPerChunk, LocalStorageSize() / 1024 / 1024,
             (unsigned long long)(CommCostPerBatchOnline() / 1024));
    return buf;
}

void SimpleBatchPianoPIR::RecordStats(double prepTime) {  // batch-pir.go:110-117
    preprocessingTime = prepTime;
    preprocessingTotal += prepTime;
    preprocessingCount += 1;
    localStorage = (uint64_t)LocalStorageSize();
    commCostPerBatchOnline = CommCostPerBatchOnline();
    SupportBatchNum = subPIR[0]->client.MaxQueryNum / QueryPerPartition;
    double DBSizeInBytes = (double)config.DBSize * (double)config.DBEntryByteNum;
    commCostPerBatchOffline = (uint64_t)(DBSizeInBytes / (double)SupportBatchNum);
}

// All sub-PIRs in ONE pm_hintgen call (the reference loops over them on ThreadNum=1 goroutines).
void SimpleBatchPianoPIR::Preprocessing() {
    FinishedBatchNum = 0;
    QueriesMadeInPartition = 0;
    auto t0 = std::chrono::steady_clock::now();
    const uint64_t E = config.DBEntrySize, PN = config.PartitionNum;
    if (resident) {
        std::vector<uint32_t> ids(PN);
        for (uint64_t i = 0; i < PN; i++) ids[i] = (uint32_t)i;
        PreprocessResident(ids, false);
        hostMode = false;   // every cache, on the host and on the device, is empty again
        RecordStats(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return;
    }
    std::vector<pm_hint_job> jobs(PN);
    std::vector<std::vector<uint64_t>> outs(PN);
    for (uint64_t i = 0; i < PN; i++) {
        PianoPIRClient &c = subPIR[i]->client;
        c.Initialization();
        outs[i].resize((c.primaryHintNum + subPIR[i]->config.SetSize * c.maxQueryPerChunk) * E);
        c.FillHintJob(subPIR[i]->server.row0, &jobs[i], outs[i].data());
    }
    check(pm_hintgen(db->h, jobs.data(), PN), "pm_hintgen");
    // replacement values for every sub-PIR in one gather (indices past a partition's end are its zero padding)
    std::vector<uint64_t> gidx, ridx;
    std::vector<uint64_t> base(PN + 1, 0);
    for (uint64_t i = 0; i < PN; i++) {
        PianoPIRClient &c = subPIR[i]->client;
        const uint64_t P = c.primaryHintNum, B = subPIR[i]->config.SetSize * c.maxQueryPerChunk;
        memcpy(c.primaryParity.data(), outs[i].data(), P * E * 8);
        memcpy(c.backupParity.data(), outs[i].data() + P * E, B * E * 8);
        c.DrawReplacementIdx(&ridx);
        for (uint64_t v : ridx) gidx.push_back(v < subPIR[i]->config.DBSize ? subPIR[i]->server.row0 + v : ~0ull);
        base[i + 1] = gidx.size();
    }
    std::vector<uint64_t> vals(gidx.size() * E);
    check(pm_gather_rows(db->h, 0, config.DBSize, gidx.data(), gidx.size(), vals.data()), "pm_gather_rows");
    for (uint64_t i = 0; i < PN; i++)
        memcpy(subPIR[i]->client.replacementVal.data(), vals.data() + base[i] * E, (base[i + 1] - base[i]) * E * 8);
    double prepTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    RecordStats(prepTime);
}

void SimpleBatchPianoPIR::DummyPreprocessing() {  // batch-pir.go:157-166
    if (resident) {
        std::vector<uint32_t> ids(config.PartitionNum);
        for (uint64_t i = 0; i < config.PartitionNum; i++) ids[i] = (uint32_t)i;
        PreprocessResident(ids, true);
        for (auto *p : subPIR) p->client.skipPrep = true;
        RecordStats(0);
        return;
    }
    for (auto *p : subPIR) p->DummyPreprocessing();
    RecordStats(0);
}

// answer every prepared sub-query with one launch, then finish them in issue order
void SimpleBatchPianoPIR::Flush(std::vector<PendingQuery> &pend, std::vector<uint64_t> &pend_part,
                                std::vector<uint64_t> &pend_global,
                                std::unordered_map<uint64_t, std::vector<uint64_t>> &responses) {
    const uint64_t E = config.DBEntrySize;
    uint64_t stride = 0, q = 0;
    for (size_t a = 0; a < pend.size(); a++)
        if (!pend[a].offsets.empty()) {
            q++;
            stride = std::max<uint64_t>(stride, pend[a].offsets.size());
        }
    std::vector<uint64_t> row0(q), nrows(q), out(q * E);
    std::vector<uint32_t> chunk(q), set(q), offs(q * stride, 0);
    uint64_t k = 0;
    for (size_t a = 0; a < pend.size(); a++) {
        if (pend[a].offsets.empty()) continue;
        PianoPIR *p = subPIR[pend_part[a]];
        row0[k] = p->server.row0;
        nrows[k] = p->config.DBSize;
        chunk[k] = (uint32_t)p->config.ChunkSize;
        set[k] = (uint32_t)p->config.SetSize;
        memcpy(&offs[k * stride], pend[a].offsets.data(), pend[a].offsets.size() * 4);
        k++;
    }
    if (q) {
        check(pm_answer_batch(db->h, row0.data(), nrows.data(), chunk.data(), set.data(), offs.data(), stride, q, out.data()),
              "pm_answer_batch");
        serverQueries += q;
        serverLaunches += 1;
    }
    k = 0;
    std::vector<uint64_t> ret;
    for (size_t a = 0; a < pend.size(); a++) {
        const uint64_t *resp = nullptr;
        if (!pend[a].offsets.empty()) resp = &out[(k++) * E];
        subPIR[pend_part[a]]->client.FinishQuery(pend[a], resp, &ret);
        if (pend[a].kind != PendingQuery::Dummy) responses[pend_global[a]] = ret;  // batch-pir.go:213 (errors store zeros)
    }
    pend.clear();
    pend_part.clear();
    pend_global.clear();
}

int SimpleBatchPianoPIR::Query(const std::vector<uint64_t> &idx, std::vector<std::vector<uint64_t>> *ret) {
    if (resident) return QueryResident(idx, ret);
    const uint64_t PN = config.PartitionNum, PS = config.PartitionSize, E = config.DBEntrySize;
    const uint64_t queryNumToMake = idx.size() / PN;
    std::vector<std::vector<uint64_t>> partitionQueries(PN);
    for (uint64_t v : idx) {
        uint64_t pi = v / PS;
        if (pi >= PN) return -1;  // Go: index out of range panic
        partitionQueries[pi].push_back(v);
    }
    std::unordered_map<uint64_t, std::vector<uint64_t>> responses;
    std::vector<PendingQuery> pend;
    std::vector<uint64_t> pend_part, pend_global;
    for (uint64_t i = 0; i < PN; i++) {
        auto &lst = partitionQueries[i];
        while (lst.size() < queryNumToMake) lst.push_back(DefaultValue);
        for (uint64_t j = 0; j < queryNumToMake; j++) {
            PianoPIR *p = subPIR[i];
            if (p->client.FinishedQueryNum == p->client.MaxQueryNum) {  // pir.go:527-530, mid-batch: settle first
                Flush(pend, pend_part, pend_global, responses);
                p->client.Preprocessing(&p->server);
            }
            pend.emplace_back();
            pend_part.push_back(i);
            if (lst[j] == DefaultValue) {
                pend_global.push_back(DefaultValue);
                p->client.PrepareQuery(0, false, &pend.back());
            } else {
                pend_global.push_back(lst[j]);
                p->client.PrepareQuery(lst[j] - i * PS, true, &pend.back());
            }
        }
    }
    Flush(pend, pend_part, pend_global, responses);
    ret->resize(idx.size());
    for (size_t i = 0; i < idx.size(); i++) {
        auto it = responses.find(idx[i]);
        if (it != responses.end()) (*ret)[i] = it->second;
        else (*ret)[i].assign(E, 0);
    }
    if (QueriesMadeInPartition >= subPIR[0]->client.MaxQueryNum - 2) {  // batch-pir.go:239-245
        Preprocessing();
    } else {
        FinishedBatchNum += idx.size() / config.BatchSize;
        QueriesMadeInPartition += queryNumToMake;
    }
    return 0;
}

int HostThreads() {
    static const int n = [] {
        int t = 0;
        if (const char *v = getenv("PM_HOST_THREADS")) t = atoi(v);
        if (t <= 0)
            if (const char *v = getenv("OMP_NUM_THREADS")) t = atoi(v);
        if (t <= 0) t = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        return std::min(t, 64);
    }();
    return n;
}

struct WorkerPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    // state: odd = a loop is open for workers (they may take items), even = closed.  The caller opens a loop only after
    // its fields are set and closes it before returning, then waits until no worker is inside: a late worker never
    // sees a half-initialised loop and never runs an item of a loop that is over.
    std::atomic<uint64_t> state{0};
    std::atomic<size_t> next{0}, done{0};
    std::atomic<int> sleeping{0}, active{0};
    std::atomic<bool> stop{false};
    size_t n = 0;
    const std::function<void(size_t)> *fn = nullptr;
    std::string error;

    void run_items() {
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) break;
            try {
                (*fn)(i);
            } catch (const std::exception &e) {
                std::lock_guard<std::mutex> lock(mu);
                if (error.empty()) error = e.what();
            }
            done.fetch_add(1, std::memory_order_release);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            // wait for the next open loop: spin ~50 us (a search step opens one every few hundred us), then sleep
            uint64_t st = state.load(std::memory_order_acquire);
            for (int spin = 0; !((st & 1) && st != seen) && spin < 20000 && !stop.load(std::memory_order_relaxed); spin++) {
                _mm_pause();
                st = state.load(std::memory_order_acquire);
            }
            if (!((st & 1) && st != seen)) {
                std::unique_lock<std::mutex> lock(mu);
                sleeping.fetch_add(1);
                cv.wait(lock, [&] {
                    const uint64_t v = state.load(std::memory_order_acquire);
                    return ((v & 1) && v != seen) || stop.load();
                });
                sleeping.fetch_sub(1);
                st = state.load(std::memory_order_acquire);
            }
            if (stop.load()) return;
            if (!((st & 1) && st != seen)) continue;
            active.fetch_add(1, std::memory_order_acq_rel);
            if (state.load(std::memory_order_acquire) == st) run_items();   // still the loop we saw open
            active.fetch_sub(1, std::memory_order_acq_rel);
            seen = st;
        }
    }
};

WorkerPool::WorkerPool(int threads) : impl(new Impl()) {
    for (int i = 1; i < threads; i++) impl->workers.emplace_back([this] { impl->worker(); });
}
WorkerPool::~WorkerPool() {
    {
        std::lock_guard<std::mutex> lock(impl->mu);
        impl->stop.store(true);
    }
    impl->cv.notify_all();
    for (auto &t : impl->workers) t.join();
    delete impl;
}
void WorkerPool::ParallelFor(size_t n, const std::function<void(size_t)> &fn) {
    if (n == 0) return;
    if (impl->workers.empty() || n <= 2) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    impl->error.clear();
    impl->fn = &fn;
    impl->n = n;
    impl->done.store(0, std::memory_order_relaxed);
    impl->next.store(0, std::memory_order_relaxed);
    const uint64_t open = impl->state.load(std::memory_order_relaxed) + 1;   // even -> odd
    {
        std::lock_guard<std::mutex> lock(impl->mu);   // pairs with the sleepers' predicate check
        impl->state.store(open, std::memory_order_release);
    }
    if (impl->sleeping.load() > 0) impl->cv.notify_all();
    impl->run_items();
    while (impl->done.load(std::memory_order_acquire) < n) _mm_pause();
    impl->state.store(open + 1, std::memory_order_release);                  // closed
    while (impl->active.load(std::memory_order_acquire) != 0) _mm_pause();   // nobody is inside any more
    if (!impl->error.empty()) throw std::runtime_error(impl->error);
}
WorkerPool &WorkerPool::Local() {
    static thread_local WorkerPool pool(HostThreads());
    return pool;
}

void EntryCache::Reserve(uint64_t entries) {
    while (slabs.size() * kPerSlab < entries) slabs.emplace_back(new uint64_t[kPerSlab * E]());   // value-initialised: pages touched now
    reserved = std::max(reserved, entries);
    if (slot.size() == 0) slot.reset(reserved);
}
const uint64_t *EntryCache::put(uint64_t idx, const uint64_t *entry) {
    uint64_t s;
    if (const uint64_t *have = slot.find(idx)) {
        s = *have;
    } else {
        s = used++;
        if (s / kPerSlab >= slabs.size()) slabs.emplace_back(new uint64_t[kPerSlab * E]);
        slot.put(idx, s);
    }
    uint64_t *dst = slabs[s / kPerSlab].get() + (s % kPerSlab) * E;
    memcpy(dst, entry, E * 8);
    return dst;
}

// ---------------------------------------------------------------------------------------------
// GPU-resident client
// ---------------------------------------------------------------------------------------------
void SimpleBatchPianoPIR::EnableResidentClient(uint32_t lanes) {
    if (resident) return;
    if (lanes == 0) lanes = 1;
    std::vector<pm_client_part> parts((size_t)config.PartitionNum * lanes);
    for (uint32_t l = 0; l < lanes; l++)
        for (uint64_t i = 0; i < config.PartitionNum; i++) {
            PianoPIR *p = subPIR[i];
            parts[(size_t)l * config.PartitionNum + i] = pm_client_part{p->server.row0, p->config.DBSize, p->config.ChunkSize, p->config.SetSize,
                                                                        p->client.primaryHintNum, p->client.maxQueryPerChunk, p->client.MaxQueryNum};
        }
    check(pm_client_create(db->h, parts.data(), parts.size(), &rclient), "pm_client_create");
    resident = true;
    ownsClient = true;
    partBase = 0;
    clientLanes = lanes;
    reserveCaches();
}

void SimpleBatchPianoPIR::AttachResidentClient(SimpleBatchPianoPIR *owner, uint32_t lane) {
    if (resident) return;
    if (!owner || !owner->resident || !owner->ownsClient || lane == 0 || lane >= owner->clientLanes)
        throw std::runtime_error("AttachResidentClient: the owner has no resident client with that lane");
    if (owner->config.PartitionNum != config.PartitionNum || owner->config.DBSize != config.DBSize ||
        owner->config.DBEntrySize != config.DBEntrySize || owner->db != db)
        throw std::runtime_error("AttachResidentClient: lanes must be clients of the same database with the same geometry");
    rclient = owner->rclient;
    resident = true;
    ownsClient = false;
    partBase = lane * (uint32_t)config.PartitionNum;
    clientLanes = owner->clientLanes;
    reserveCaches();
}

// a sub-PIR caches at most MaxQueryNum entries between two preprocessings (pir.go:468 runs once per successful query)
void SimpleBatchPianoPIR::reserveCaches() {
    for (auto *p : subPIR) {
        p->client.localCache.Init(config.DBEntrySize);
        p->client.localCache.Reserve(p->client.MaxQueryNum);
    }
}

// Initialization + Preprocessing of the listed sub-PIRs on the device (keys and seeds derived as in the host path)
void SimpleBatchPianoPIR::PreprocessResident(const std::vector<uint32_t> &ids, bool skipPrep) {
    std::vector<uint32_t> rk(ids.size() * 44);
    std::vector<uint64_t> seeds(ids.size());
    std::vector<uint8_t> keys(ids.size() * 16);
    for (size_t a = 0; a < ids.size(); a++) {
        PianoPIRClient &c = subPIR[ids[a]]->client;
        c.FinishedQueryNum = 0;
        c.prepEpoch = c.keyEpoch++;   // a fresh key per preprocessing, as in Initialization()
        c.masterKey = DeriveKey(c.keySeed, c.prepEpoch, c.keyParts, c.keyIndex, c.keySeedHi);
        memcpy(&keys[a * 16], c.masterKey.b, 16);
        c.localCache.Init(config.DBEntrySize);
        c.pendingCached.clear();
        seeds[a] = Mix64(c.replSeed, c.prepEpoch * c.keyParts + c.keyIndex);
    }
    check(pm_expand_key_batch(keys.data(), ids.size(), rk.data()), "pm_expand_key_batch");  // GetLongKey for every sub-PIR
    for (size_t a = 0; a < ids.size(); a++) subPIR[ids[a]]->client.longKey.assign(rk.begin() + a * 44, rk.begin() + (a + 1) * 44);
    std::vector<uint32_t> devIds(ids);
    for (auto &v : devIds) v += partBase;
    check(pm_client_preprocess(rclient, devIds.data(), devIds.size(), rk.data(), seeds.data(), skipPrep ? 1 : 0), "pm_client_preprocess");
}

void SimpleBatchPianoPIR::SyncTablesFromDevice(uint64_t i) {
    if (!resident) return;
    PianoPIRClient &c = subPIR[i]->client;
    const uint64_t E = config.DBEntrySize, P = c.primaryHintNum, B = subPIR[i]->config.SetSize * c.maxQueryPerChunk;
    c.primaryShortTag.resize(P); c.primaryParity.resize(P * E); c.primaryProgramPoint.resize(P);
    c.replacementIdx.resize(B); c.replacementVal.resize(B * E); c.backupShortTag.resize(B); c.backupParity.resize(B * E);
    c.QueryHistogram.resize(subPIR[i]->config.SetSize);
    std::vector<uint64_t> *tabs[8] = {&c.primaryShortTag, &c.primaryParity, &c.primaryProgramPoint, &c.replacementIdx,
                                      &c.replacementVal, &c.backupShortTag, &c.backupParity, &c.QueryHistogram};
    for (int t = 0; t < 8; t++) check(pm_client_download(rclient, partBase + (uint32_t)i, t, tabs[t]->data(), tabs[t]->size()), "pm_client_download");
    uint64_t fin = 0;
    check(pm_client_download(rclient, partBase + (uint32_t)i, 8, &fin, 1), "pm_client_download");
    c.FinishedQueryNum = fin;
}

int SimpleBatchPianoPIR::QueryResident(const std::vector<uint64_t> &idx, std::vector<std::vector<uint64_t>> *ret) {
    const uint64_t E = config.DBEntrySize;
    std::vector<uint64_t> flat(idx.size() * E);
    int rc = QueryFlat(idx.data(), idx.size(), flat.data(), nullptr, 0, nullptr);
    if (rc != 0) return rc;
    ret->resize(idx.size());
    for (size_t i = 0; i < idx.size(); i++) (*ret)[i].assign(flat.begin() + i * E, flat.begin() + (i + 1) * E);
    return 0;
}

// ---- the pieces of one resident Query call (batch-pir.go:170-248 over pm_client_*) ----
// bucket the indices by partition (batch-pir.go:177-187) and reset the per-call scratch
void SimpleBatchPianoPIR::beginCall(const uint64_t *idx, size_t n, bool *bad) {
    const uint64_t PN = config.PartitionNum, PS = config.PartitionSize, E = config.DBEntrySize;
    *bad = false;
    wsLists.resize(PN);
    for (auto &l : wsLists) l.clear();
    for (size_t i = 0; i < n; i++) {
        uint64_t pi = idx[i] / PS;
        if (pi >= PN) { *bad = true; return; }
        wsLists[pi].push_back(idx[i]);
    }
    const uint64_t queryNumToMake = n / PN;
    for (auto &l : wsLists) while (l.size() < queryNumToMake) l.push_back(DefaultValue);
    wsResponses.reset(n);
    wsRespList.clear();
    wsPend.clear();
    wsQueries.clear();
    wsZero.assign(E, 0);
    wsPendingReal.assign(PN, 0);
    wsCached.resize(n * E);   // sized up front: the pointers handed out below must stay valid for the whole call
    wsCachedUsed = 0;
}

// one sub-query of partition `part` (batch-pir.go:189-216 / pir.go:354-383): a dummy, a local-cache hit, or a record for the GPU
void SimpleBatchPianoPIR::pushRecord(uint64_t part, uint64_t globalIdx) {
    PianoPIR *p = subPIR[part];
    PianoPIRClient &c = p->client;
    if (globalIdx == DefaultValue) {
        wsQueries.push_back(pm_client_query{partBase + (uint32_t)part, 0, 0, c.dummySeed, c.dummyCtr});
        c.dummyCtr += p->config.SetSize;
        wsPend.push_back(PendRec{part, DefaultValue, 0, 0, (int64_t)wsQueries.size() - 1});
        return;
    }
    const uint64_t local = globalIdx - part * config.PartitionSize;
    bool cached = c.localCache.has(local);
    for (size_t k = 0; !cached && k < c.pendingCached.size(); k++) cached = c.pendingCached[k] == local;
    if (cached) {
        wsPend.push_back(PendRec{part, globalIdx, local, 2, -1});
    } else {
        wsQueries.push_back(pm_client_query{partBase + (uint32_t)part, 1, local, 0, 0});
        c.pendingCached.push_back(local);
        wsPendingReal[part] += 1;
        wsPend.push_back(PendRec{part, globalIdx, local, 1, (int64_t)wsQueries.size() - 1});
    }
}

// book the answers of the records wsPend[pbase..]: res / status / dist are indexed by the record's position in wsQueries
void SimpleBatchPianoPIR::settle(size_t pbase, const uint64_t *res, const int32_t *status, const float *dist) {
    const uint64_t E = config.DBEntrySize;
    const float kNaN = std::nanf("");
    for (size_t a = pbase; a < wsPend.size(); a++) {
        const PendRec &pd = wsPend[a];
        PianoPIRClient &c = subPIR[pd.part]->client;
        if (pd.kind == 0) { serverQueries += 1; continue; }
        if (pd.kind == 2) {  // served from the local cache (pir.go:381-383); an earlier failure of the same index repeats as zeros
            // copied out of the cache: a budget-triggered re-preprocessing later in this call clears the cache and
            // reuses its slabs, and out_ptrs callers keep these pointers until the lane's next call
            const uint64_t *e = c.localCache.find(pd.local);
            if (e) {
                uint64_t *dst = wsCached.data() + (wsCachedUsed++) * E;
                memcpy(dst, e, E * 8);
                e = dst;
            }
            wsResponses.put(pd.global, wsRespList.size());
            wsRespList.push_back(Resp{e ? e : wsZero.data(), kNaN});
            continue;
        }
        const uint64_t *r = res + (size_t)pd.qpos * E;
        if (status[pd.qpos] == 0) {
            serverQueries += 1;
            c.FinishedQueryNum += 1;
            c.localCache.put(pd.local, r);
        }
        for (size_t k = 0; k < c.pendingCached.size(); k++)
            if (c.pendingCached[k] == pd.local) { c.pendingCached.erase(c.pendingCached.begin() + (long)k); break; }
        wsResponses.put(pd.global, wsRespList.size());
        wsRespList.push_back(Resp{r, dist ? dist[pd.qpos] : kNaN});
    }
    std::fill(wsPendingReal.begin(), wsPendingReal.end(), 0);
}

// responses keyed by global index, zero rows for misses (batch-pir.go:218-237), then the batch accounting (:239-245)
bool SimpleBatchPianoPIR::finishCall(const uint64_t *idx, size_t n, uint64_t *out, const uint64_t **out_ptrs, float *dists) {
    const uint64_t E = config.DBEntrySize;
    const float kNaN = std::nanf("");
    for (size_t i = 0; i < n; i++) {
        const uint64_t *pos = wsResponses.find(idx[i]);
        const bool have = pos != nullptr;
        const Resp *r = have ? &wsRespList[*pos] : nullptr;
        if (out_ptrs) out_ptrs[i] = have ? r->entry : wsZero.data();
        else if (have) memcpy(out + i * E, r->entry, E * 8);
        else memset(out + i * E, 0, E * 8);
        if (dists) dists[i] = have ? r->dist : kNaN;
    }
    if (QueriesMadeInPartition >= subPIR[0]->client.MaxQueryNum - 2) return true;
    FinishedBatchNum += n / config.BatchSize;
    QueriesMadeInPartition += n / config.PartitionNum;
    return false;
}

// pir.go:527-530 re-preprocesses a sub-PIR whose query budget is used up; that needs the exact FinishedQueryNum, i.e.
// the pending queries must have run.  True when this call could reach that point for some sub-PIR.
bool SimpleBatchPianoPIR::mayFlushInside(size_t n) const {
    const uint64_t per = n / config.PartitionNum;
    for (auto *p : subPIR)
        if (p->client.FinishedQueryNum + per >= p->client.MaxQueryNum) return true;
    return false;
}

// Flat form of Query for the resident client: out is [n][DBEntrySize].  With query_vec != nullptr it also returns
// dists[i] = L2Dist(vector part of out[i], query_vec) from the same GPU call (NaN where the entry came from the
// local cache and no distance was computed).  Scratch vectors are members: no allocation on the steady-state path.
int SimpleBatchPianoPIR::QueryFlat(const uint64_t *idx, size_t n, uint64_t *out, const float *query_vec, uint64_t dim, float *dists) {
    const uint64_t PN = config.PartitionNum, E = config.DBEntrySize;
    if (!resident) {
        std::vector<uint64_t> v(idx, idx + n);
        std::vector<std::vector<uint64_t>> r;
        int rc = Query(v, &r);
        if (rc != 0) return rc;
        for (size_t i = 0; i < n; i++) memcpy(out + i * E, r[i].data(), E * 8);
        if (dists) for (size_t i = 0; i < n; i++) dists[i] = std::nanf("");
        return 0;
    }
    struct Timer {
        double &acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~Timer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    } timer{profQueryTotal};
    profQueryCalls += 1;
    bool bad = false;
    beginCall(idx, n, &bad);
    if (bad) return -1;
    const uint64_t queryNumToMake = n / PN;
    wsOut.resize(n * E);
    wsStatus.resize(n);
    wsDist.resize(n);
    size_t qbase = 0, pbase = 0;   // queries / pending records already settled by an earlier flush of this call

    auto flush = [&]() {
        const size_t cnt = wsQueries.size() - qbase;
        if (cnt) {
            auto tg = std::chrono::steady_clock::now();
            if (query_vec)
                check(pm_client_query_batch_l2(rclient, wsQueries.data() + qbase, cnt, wsOut.data() + qbase * E, wsStatus.data() + qbase,
                                               query_vec, dim, wsDist.data() + qbase), "pm_client_query_batch_l2");
            else
                check(pm_client_query_batch(rclient, wsQueries.data() + qbase, cnt, wsOut.data() + qbase * E, wsStatus.data() + qbase),
                      "pm_client_query_batch");
            profGpuCall += std::chrono::duration<double>(std::chrono::steady_clock::now() - tg).count();
            serverLaunches += 1;
        }
        settle(pbase, wsOut.data(), wsStatus.data(), query_vec ? wsDist.data() : nullptr);
        qbase = wsQueries.size();
        pbase = wsPend.size();
    };

    for (uint64_t i = 0; i < PN; i++) {
        PianoPIRClient &c = subPIR[i]->client;
        for (uint64_t j = 0; j < queryNumToMake; j++) {
            // pir.go:527-530 needs the exact FinishedQueryNum, which is only known after the pending queries ran
            if (c.FinishedQueryNum + wsPendingReal[i] >= c.MaxQueryNum) {
                flush();
                if (c.FinishedQueryNum == c.MaxQueryNum) PreprocessResident({(uint32_t)i}, c.skipPrep);
            }
            pushRecord(i, wsLists[i][j]);
        }
    }
    flush();
    if (finishCall(idx, n, out, nullptr, dists)) Preprocessing();
    return 0;
}

// Several lanes of one pm_client in one device call.  Host work per lane (bucketing, cache checks, booking the answers)
// runs in parallel over the lanes; the lanes share nothing on the host.
int SimpleBatchPianoPIR::QueryFlatGroup(std::vector<GroupCall> &calls, uint64_t dim) {
    if (calls.empty()) return 0;
    static const bool prof = getenv("PM_HOST_PROFILE") != nullptr;
    static thread_local double tBuild = 0, tDev = 0, tSettle = 0;
    static thread_local uint64_t nCalls = 0;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](double &acc) {
        auto t = std::chrono::steady_clock::now();
        acc += std::chrono::duration<double>(t - tp0).count();
        tp0 = t;
    };
    const size_t L = calls.size();
    pm_client *client = calls[0].pir->rclient;
    const uint64_t E = calls[0].pir->config.DBEntrySize;
    std::vector<char> grouped(L, 0);
    for (size_t l = 0; l < L; l++) {
        SimpleBatchPianoPIR *p = calls[l].pir;
        calls[l].rc = 0;
        grouped[l] = p->resident && p->rclient == client && client != nullptr && !p->mayFlushInside(calls[l].n) && p->config.DBEntrySize == E;
    }
    // lanes that cannot take part (not resident, another client, a sub-PIR about to exhaust its budget) run on their own
    for (size_t l = 0; l < L; l++) {
        if (grouped[l]) continue;
        GroupCall &gc = calls[l];
        uint64_t *dst = gc.out;
        if (!dst) {   // pointer form: the lane's own scratch holds the copies
            gc.pir->wsSolo.resize(gc.n * E);
            dst = gc.pir->wsSolo.data();
            for (size_t i = 0; i < gc.n; i++) gc.out_ptrs[i] = dst + i * E;
        }
        gc.rc = gc.pir->QueryFlat(gc.idx, gc.n, dst, gc.query_vec, dim, gc.dists);
    }
    std::vector<size_t> base(L + 1, 0);
    WorkerPool &pool = WorkerPool::Local();
    pool.ParallelFor(L, [&](size_t l) {
        if (!grouped[l]) return;
        SimpleBatchPianoPIR *p = calls[l].pir;
        bool bad = false;
        p->profQueryCalls += 1;
        p->beginCall(calls[l].idx, calls[l].n, &bad);
        if (bad) { calls[l].rc = -1; p->wsQueries.clear(); p->wsPend.clear(); return; }
        const uint64_t per = calls[l].n / p->config.PartitionNum;
        for (uint64_t i = 0; i < p->config.PartitionNum; i++)
            for (uint64_t j = 0; j < per; j++) p->pushRecord(i, p->wsLists[i][j]);
    });
    for (size_t l = 0; l < L; l++) base[l + 1] = base[l] + (grouped[l] ? calls[l].pir->wsQueries.size() : 0);
    const size_t total = base[L];
    // group scratch lives in the first grouped lane's members
    SimpleBatchPianoPIR *host = nullptr;
    for (size_t l = 0; l < L && !host; l++) if (grouped[l]) host = calls[l].pir;
    if (!host) return 0;
    static thread_local std::vector<pm_client_query> gq;
    static thread_local std::vector<uint32_t> gvec;
    static thread_local std::vector<float> gqv;
    gq.resize(total);
    gvec.resize(total);
    gqv.assign(L * dim, 0.f);
    bool anyVec = false;
    for (size_t l = 0; l < L; l++) {
        if (!grouped[l]) continue;
        SimpleBatchPianoPIR *p = calls[l].pir;
        std::copy(p->wsQueries.begin(), p->wsQueries.end(), gq.begin() + (long)base[l]);
        std::fill(gvec.begin() + (long)base[l], gvec.begin() + (long)base[l + 1], (uint32_t)l);
        if (calls[l].query_vec) { memcpy(&gqv[l * dim], calls[l].query_vec, dim * 4); anyVec = true; }
    }
    lap(tBuild);
    if (host->groupBufWords < total * E) {   // page-locked: the answers arrive by DMA, no host-side copy in the C-ABI
        if (host->groupBuf) pm_host_free(host->groupBuf);
        host->groupBuf = nullptr;
        host->groupBufWords = 0;
        void *pbuf = nullptr;
        check(pm_host_alloc(&pbuf, (total * E + total * E / 2) * 8), "pm_host_alloc");
        host->groupBuf = (uint64_t *)pbuf;
        host->groupBufWords = total * E + total * E / 2;
    }
    host->wsStatus.resize(std::max(host->wsStatus.size(), total));
    host->wsDist.resize(std::max(host->wsDist.size(), total));
    if (total) {
        auto tg = std::chrono::steady_clock::now();
        if (anyVec && dim)
            check(pm_client_query_batch_l2m(client, gq.data(), total, host->groupBuf, host->wsStatus.data(), gqv.data(), L, gvec.data(), dim,
                                            host->wsDist.data()), "pm_client_query_batch_l2m");
        else
            check(pm_client_query_batch(client, gq.data(), total, host->groupBuf, host->wsStatus.data()), "pm_client_query_batch");
        host->profGpuCall += std::chrono::duration<double>(std::chrono::steady_clock::now() - tg).count();
        host->serverLaunches += 1;
    }
    lap(tDev);
    const uint64_t *gout = host->groupBuf;
    const int32_t *gst = host->wsStatus.data();
    const float *gdist = host->wsDist.data();
    std::vector<char> due(L, 0);
    pool.ParallelFor(L, [&](size_t l) {
        if (!grouped[l] || calls[l].rc != 0) return;
        SimpleBatchPianoPIR *p = calls[l].pir;
        p->settle(0, gout + base[l] * E, gst + base[l], (anyVec && dim && calls[l].query_vec) ? gdist + base[l] : nullptr);
        due[l] = p->finishCall(calls[l].idx, calls[l].n, calls[l].out, calls[l].out ? nullptr : calls[l].out_ptrs, calls[l].dists) ? 1 : 0;
    });
    lap(tSettle);
    if (prof && (++nCalls % 200) == 0)
        fprintf(stderr, "[group profile] %llu calls, us per call: build %.1f | device call %.1f | settle %.1f\n", (unsigned long long)nCalls,
                tBuild / nCalls * 1e6, tDev / nCalls * 1e6, tSettle / nCalls * 1e6);
    for (size_t l = 0; l < L; l++)
        if (due[l]) calls[l].pir->Preprocessing();   // batch-pir.go:239-245; device calls of one pm_client are serialised anyway
    for (size_t l = 0; l < L; l++)
        if (calls[l].rc != 0) return calls[l].rc;
    return 0;
}

// ---- host side of the device-resident search path ----
bool SimpleBatchPianoPIR::DeviceFetchAccounting(size_t n) {
    profQueryCalls += 1;
    serverLaunches += 1;
    if (QueriesMadeInPartition >= subPIR[0]->client.MaxQueryNum - 2) return true;
    FinishedBatchNum += n / config.BatchSize;
    QueriesMadeInPartition += n / config.PartitionNum;
    return false;
}
void SimpleBatchPianoPIR::AbsorbDeviceRound(const uint64_t *finished, uint64_t serverQ) {
    for (uint64_t i = 0; i < config.PartitionNum; i++) subPIR[i]->client.FinishedQueryNum = finished[i];
    serverQueries += serverQ;
}
bool SimpleBatchPianoPIR::DeviceRoundIsSafe(uint64_t maxStep, size_t n) const {
    const uint64_t per = n / config.PartitionNum;
    for (auto *p : subPIR)
        if (p->client.FinishedQueryNum + maxStep * per >= p->client.MaxQueryNum) return false;
    return true;
}
void SimpleBatchPianoPIR::EnterHostMode(pm_search *s) {
    if (hostMode) return;
    const uint64_t E = config.DBEntrySize, PN = config.PartitionNum;
    std::vector<uint64_t> idx, entries;
    for (uint64_t i = 0; i < PN; i++) {
        PianoPIRClient &c = subPIR[i]->client;
        uint64_t cnt = 0;
        check(pm_search_cache_download(s, partBase / (uint32_t)PN, (uint32_t)i, nullptr, nullptr, 0, &cnt), "pm_search_cache_download");
        c.localCache.Init(E);
        c.localCache.Reserve(c.MaxQueryNum);
        c.pendingCached.clear();
        if (cnt == 0) continue;
        idx.resize(cnt);
        entries.resize(cnt * E);
        check(pm_search_cache_download(s, partBase / (uint32_t)PN, (uint32_t)i, idx.data(), entries.data(), cnt, &cnt), "pm_search_cache_download");
        for (uint64_t k = 0; k < cnt; k++) c.localCache.put(idx[k], &entries[k * E]);
    }
    hostMode = true;
}

double SimpleBatchPianoPIR::LocalStorageSize() const {
    double r = 0;
    for (auto *p : subPIR) r += p->LocalStorageSize();
    return r;
}
uint64_t SimpleBatchPianoPIR::CommCostPerBatchOnline() const {
    double r = 0;
    for (auto *p : subPIR) r += p->CommCostPerQuery() * (double)QueryPerPartition;
    return (uint64_t)r;
}

}  // namespace pianopir
