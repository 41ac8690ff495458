// GPU-resident PianoPIR client: the hint tables of every sub-PIR of one
// SimpleBatchPianoPIR live in HBM, so Preprocessing() never ships parities to the host and the online
// Query's first-match hint search (pianopir/pir.go:405-414, up to primaryHintNum PRF evaluations per
// query on the CPU) runs as a block-wide parallel search.  Included at the end of pm_pir.cu (same
// translation unit as the AES table in constant memory).
//
// One query call = three launches on one stream (the distances of pm_client_query_batch_l2[m] ride in the third):
//   client_prepare_kernel   one CTA per sub-PIR walks its queries IN ORDER: budget checks, hint search,
//                           set expansion, programmed-point and replacement patching (pir.go:386-447) and
//                           the response-independent half of the refresh (tags, program point, counters)
//   answer_kernel           the server's PrivateQuery for every sub-query of the call (pir.go:65-88)
//   client_finish_kernel    response xor replacement xor parity, parity refresh from the backup hint
//                           (pir.go:451-463), again in order per sub-PIR
// The split at the server call is legal because nothing the prepare step reads depends on a response
// value; results and the final client state are bit-identical to the sequential reference
// (tests/test_resident_client_gpu.py compares every table with the oracle's).
#pragma once

namespace pm {

constexpr uint64_t kDefaultProgramPoint = 0x7fffffffull;  // pir.go:15
constexpr int CL_THREADS = 512;   // widest CTA of client_prepare_kernel; calls with many parts use 256 (see client_query_impl)

struct ClientPartDev {
    uint32_t rk[44];
    uint64_t row0, n_rows, chunk_size, set_size, n_primary, backup_group, max_query_num;
    uint32_t chunk_mask, chunk_shift;
    uint64_t *tags, *pp, *parity;              // [P], [P], [P][E]
    uint64_t *btags, *bparity, *ridx, *rval;   // [S*M], [S*M][E], [S*M], [S*M][E]
    uint64_t *hist;                            // [S]
    uint64_t *finished;                        // [1]
    uint16_t *poff;                            // [S][P] offset index of the primary hints (nullptr if chunk_size > 65536)
    uint16_t *roff;                            // [P+B][spad] the offsets of EVERY hint, row-major, as the hint kernel emits them:
    uint32_t spad;                             // the promoted backup hint's row is copied from here (no AES online); spad = (S+7)&~7
    uint32_t *pp32;                            // [P] program points once more as u32 (all values < 2^31), 16-byte aligned: the
                                               // prepare kernel mirrors them in shared memory with a few vector loads
};

__device__ __forceinline__ uint64_t mix64_dev(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Initialization (pir.go:203-255) + replacement index draw (pir.go:345-347), one CTA column per part
__global__ void client_init_kernel(const ClientPartDev *parts, const uint32_t *part_ids, const uint64_t *repl_seed, int skip_prep) {
    const ClientPartDev &D = parts[part_ids[blockIdx.y]];
    const uint64_t P = D.n_primary, B = D.set_size * D.backup_group, M = D.backup_group, seed = repl_seed[blockIdx.y];
    const uint64_t n = P > B ? P : B;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i < P) {
            D.tags[i] = i;
            D.pp[i] = kDefaultProgramPoint;
            D.pp32[i] = (uint32_t)kDefaultProgramPoint;
        }
        if (i < B) {
            D.btags[i] = P + i;
            const uint64_t c = i / M;
            // i == c*M + j; DummyPreprocessing leaves the Initialization value (pir.go:245)
            D.ridx[i] = skip_prep ? kDefaultProgramPoint : (mix64_dev(seed, i) & (D.chunk_size - 1)) + c * D.chunk_size;
        }
        if (i < D.set_size) D.hist[i] = 0;
        if (i == 0) *D.finished = 0;
    }
}

// The queries of sub-PIR `part`, in array order (= processing order): positions part_start[part] .. part_start[part+1]
// of part_items.  The host builds these lists (a counting sort of q records) -- an earlier version let every CTA scan
// the whole query array for its own entries, which at 3072 queries and 512 CTAs was most of the kernels' time.
constexpr uint32_t CL_MAX_LIST = 2048;   // queries of one sub-PIR per call
// Device-built calls (pm_search.cuh) use a fixed layout instead: CTA b owns records [b*fixed_per, (b+1)*fixed_per).
__device__ __forceinline__ uint32_t client_load_list(const uint32_t *part_start, const uint32_t *part_items, uint32_t part,
                                                      uint32_t *s_list, uint32_t cap, uint32_t fixed_per = 0) {
    if (fixed_per) {
        for (uint32_t i = threadIdx.x; i < fixed_per; i += blockDim.x) s_list[i] = blockIdx.x * fixed_per + i;
        __syncthreads();
        return fixed_per;
    }
    const uint32_t b = part_start[part], n = min(part_start[part + 1] - b, cap);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_list[i] = part_items[b + i];
    __syncthreads();
    return n;
}

struct RkOfPtr {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator[](int i) const { return p[i]; }
};

// Offset index.  poff[c][i] = PRF(tag_i, c) & (ChunkSize-1) for the primary hints: with it the online hint search
// (pir.go:405-414, up to primaryHintNum AES evaluations per query) is a scan of one 2*P-byte column and the set expansion
// (pir.go:424-427) a strided read of one row.  roff[h][c] holds the same for EVERY hint in hint-number order, row-major:
// the hint kernel evaluates all those PRFs anyway and writes them as a side output (pm_hint_job.offsets_out), so building
// the index costs a transpose, and a promoted backup hint's row (pir.go:460-463) is a copy -- the online path runs no AES.
// DummyPreprocessing (no hint kernel) fills roff here instead.
__global__ void __launch_bounds__(256) client_fill_roff_kernel(const ClientPartDev *parts, const uint32_t *part_ids) {
    extern __shared__ uint32_t smem[];
    __shared__ uint32_t s_rk[44];
    const ClientPartDev &D = parts[part_ids[blockIdx.y]];
    aes_tab_fill<1>(smem, c_te0);
    if (threadIdx.x < 44) s_rk[threadIdx.x] = D.rk[threadIdx.x];
    __syncthreads();
    if (!D.roff) return;
    const AesTab<1> T{smem + (threadIdx.x & 31)};
    const RkOfPtr R{s_rk};
    const uint64_t H = D.n_primary + D.set_size * D.backup_group;
    const uint32_t S = (uint32_t)D.set_size, cmask = D.chunk_mask;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < H; i += (uint64_t)gridDim.x * blockDim.x) {
        const PrfTagPart g = prf_tag_part(T, R, i);      // Initialization numbering: tag == hint number
#pragma unroll 2
        for (uint32_t c = 0; c < S; c++) D.roff[i * D.spad + c] = (uint16_t)(prf_low<1, 2>(T, R, g, c) & cmask);
    }
}
// poff[c][i] = roff[i][c] for the primary hints (tiles of 32 x 32 through shared memory)
__global__ void __launch_bounds__(256) client_transpose_kernel(const ClientPartDev *parts, const uint32_t *part_ids) {
    __shared__ uint16_t tile[32][34];
    const ClientPartDev &D = parts[part_ids[blockIdx.z]];
    if (!D.poff) return;
    const uint64_t P = D.n_primary;
    const uint32_t S = (uint32_t)D.set_size, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint64_t i0 = (uint64_t)blockIdx.x * 32; i0 < P; i0 += (uint64_t)gridDim.x * 32)
        for (uint32_t c0 = blockIdx.y * 32; c0 < S; c0 += gridDim.y * 32) {
            for (uint32_t r = ty; r < 32; r += 8)
                if (i0 + r < P && c0 + tx < S) tile[r][tx] = D.roff[(i0 + r) * D.spad + c0 + tx];
            __syncthreads();
            for (uint32_t r = ty; r < 32; r += 8)
                if (c0 + r < S && i0 + tx < P) D.poff[(uint64_t)(c0 + r) * P + i0 + tx] = tile[tx][r];
            __syncthreads();
        }
}

struct ClientQueryDev {  // mirrors pm_client_query
    uint32_t part, kind;
    uint64_t idx, dummy_seed, dummy_ctr;
};
struct ClientMeta {
    uint64_t hit, slot;
    int32_t status;  // 0 ok, 2 budget, 3 too many in chunk, 4 no hit; -1 dummy
    int32_t pad;
};

// Per query the critical path is two global round trips and two barriers:
//   A  the query's column of the offset index (the first-match search of pir.go:405-414 without any AES; program points
//      mirrored in shared memory) -- while one thread fetches the budget counters and, behind them, the replacement index
//      and the backup tag of the slot this query will consume;
//   B  the hit hint's row of the index (= the set expansion of pir.go:424-427) together with the promoted backup hint's
//      row; every thread patches its own chunk (programmed point, replacement) and stores the offset for the server and
//      the new index entry, thread 0 books the response-independent half of the refresh.
// Instances without the 16-bit index (chunk_size > 65536) evaluate the PRF instead.
__global__ void __launch_bounds__(CL_THREADS) client_prepare_kernel(const ClientPartDev *parts, const ClientQueryDev *queries,
                                                                    const uint32_t *part_start, const uint32_t *part_items,
                                                                    uint32_t q, uint32_t stride, uint32_t *offsets,
                                                                    ClientMeta *meta, uint64_t *a_row0, uint64_t *a_nrows,
                                                                    uint32_t *a_chunk, uint32_t *a_set, uint32_t mirror,
                                                                    const uint32_t *part_map, uint32_t fixed_per) {
    extern __shared__ uint32_t smem[];
    const uint32_t NT = blockDim.x;
    uint32_t *s_tab = smem;                           // compact Te0 (4 KB): few PRFs here, the shared memory buys occupancy
    uint32_t *s_rk = smem + aes_tab_words<8>();       // 44 round-key words
    uint32_t *s_offs = s_rk + 64;                     // [stride] (PRF path only)
    uint32_t *s_list = s_offs + stride;               // [CL_MAX_LIST]
    uint32_t *s_pp = s_list + CL_MAX_LIST;            // [P] program points (offset index in use and `mirror`: they fit in shared
                                                      // memory; pir_test's 2^20-row instance has P = 59392 and reads pp32 instead)
    __shared__ uint32_t s_hit;
    __shared__ int s_status;
    __shared__ uint64_t s_ingroup, s_fin, s_ridx, s_btag;
    const uint32_t part = part_map ? part_map[blockIdx.x] : blockIdx.x;   // CTA b works for sub-PIR part_map[b] (device-built calls)
    const ClientPartDev &D = parts[part];
    const uint32_t S = (uint32_t)D.set_size, cmask = D.chunk_mask;
    const uint64_t C = D.chunk_size, M = D.backup_group, P = D.n_primary;
    if (!fixed_per && part_start[part] == part_start[part + 1]) return;
    const uint32_t n_mine = client_load_list(part_start, part_items, part, s_list, CL_MAX_LIST, fixed_per);
    const bool indexed = D.poff != nullptr;
    if (!indexed) {   // only instances without the 16-bit offset index evaluate the PRF online
        aes_tab_fill<8>(s_tab, c_te0);
        if (threadIdx.x < 44) s_rk[threadIdx.x] = D.rk[threadIdx.x];
    }
    if (indexed && mirror) {   // values < 2^31 or 0x7fffffff; s_pp is 16-byte aligned (every region before it is a multiple of 4 words)
        const uint4 *src = reinterpret_cast<const uint4 *>(D.pp32);
        uint4 *dst = reinterpret_cast<uint4 *>(s_pp);
        for (uint64_t i = threadIdx.x; i < P / 4; i += NT) dst[i] = __ldcg(src + i);
        for (uint64_t i = (P & ~3ull) + threadIdx.x; i < P; i += NT) s_pp[i] = D.pp32[i];
    }
    if (threadIdx.x == 0) { s_fin = *D.finished; s_hit = 0xffffffffu; }   // only this CTA changes the part's counters during the call
    // with the offset index s_offs is free: it mirrors the per-chunk query counters, so that the slot a query consumes is
    // known at once and its replacement index / backup tag are fetched alongside the column scan, not behind a counter load
    if (indexed)
        for (uint32_t c = threadIdx.x; c < S; c += NT) s_offs[c] = (uint32_t)D.hist[c];
    __syncthreads();
    const AesTab<8> T{s_tab + (threadIdx.x & 3)};
    const RkOfPtr R{s_rk};

    for (uint32_t k = 0; k < n_mine; k++) {
        const uint32_t t = s_list[k];
        const ClientQueryDev Q = queries[t];
        if (threadIdx.x == 0) {
            a_row0[t] = D.row0; a_nrows[t] = D.n_rows; a_chunk[t] = (uint32_t)C; a_set[t] = S;
        }
        if (Q.kind >= 2) {  // device-built call: a slot served from the client's cache or by an earlier record -- no server query
            if (threadIdx.x == 0) { meta[t] = ClientMeta{0, 0, -2, 0}; a_set[t] = 0; }
            continue;
        }
        if (Q.kind == 0) {  // dummy query: SetSize random offsets (pir.go:363-371)
            for (uint32_t c = threadIdx.x; c < S; c += NT)
                offsets[(uint64_t)t * stride + c] = (uint32_t)(mix64_dev(Q.dummy_seed, Q.dummy_ctr + c) & cmask);
            if (threadIdx.x == 0) meta[t] = ClientMeta{0, 0, -1, 0};
            continue;
        }
        const uint64_t chunkId = Q.idx / C, offset = Q.idx % C;
        // ---- phase A: budget counters + first-match search (pir.go:386-414); s_hit was reset behind the last barrier ----
        if (threadIdx.x == 32 % NT) {   // a lane of another warp than the one that is busiest in the scan tail
            const uint64_t fin = s_fin, h = indexed ? (uint64_t)s_offs[chunkId] : D.hist[chunkId];
            const int st = fin >= D.max_query_num ? 2 : (h >= M ? 3 : 0);   // pir.go:386-391, 396-400
            s_status = st;
            s_ingroup = h;
            if (st == 0) {   // the slot this query consumes: replacement index and backup tag (pir.go:436-439, 460)
                const uint64_t slot = chunkId * M + h;
                s_ridx = D.ridx[slot];
                s_btag = D.btags[slot];
            }
        }
        uint32_t hit = 0xffffffffu;
        if (indexed) {  // scan one column of the offset index: no AES
            // all loads of the column are issued before the first compare (a compare-and-branch per load would
            // serialise one L2 round trip per element): 16-byte vectors of 8 offsets, up to 4 per thread per pass
            const uint16_t *col = D.poff + chunkId * P;
            auto consider = [&](uint64_t i, uint32_t v) {
                if (v == (uint32_t)offset) {
                    const uint32_t pp = mirror ? s_pp[i] : __ldcg(D.pp32 + i);
                    if (pp == (uint32_t)kDefaultProgramPoint || pp / C != chunkId) atomicMin(&s_hit, (uint32_t)i);
                }
            };
            if ((P & 7) == 0) {
                const uint4 *col4 = reinterpret_cast<const uint4 *>(col);
                const uint64_t nv = P / 8;
                const uint32_t off2 = (uint32_t)offset | ((uint32_t)offset << 16);
                for (uint64_t v0 = 0; v0 < nv; v0 += 4 * NT) {
                    uint4 w[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint64_t vi = v0 + u * NT + threadIdx.x;
                        w[u] = vi < nv ? __ldcg(col4 + vi) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint64_t vi = v0 + u * NT + threadIdx.x;
                        if (vi >= nv) continue;
                        const uint32_t x[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
                        // two 16-bit offsets per compare (SIMD-in-word); a match is rare (P / ChunkSize per column)
                        const uint32_t m0 = __vcmpeq2(x[0], off2), m1 = __vcmpeq2(x[1], off2), m2 = __vcmpeq2(x[2], off2), m3 = __vcmpeq2(x[3], off2);
                        if ((m0 | m1 | m2 | m3) == 0) continue;
                        const uint32_t m[4] = {m0, m1, m2, m3};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (m[e] & 0xffffu) consider(vi * 8 + 2 * e, (uint32_t)offset);
                            if (m[e] >> 16) consider(vi * 8 + 2 * e + 1, (uint32_t)offset);
                        }
                    }
                }
            } else {
                for (uint64_t i0 = 0; i0 < P; i0 += 8 * NT) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const uint64_t i = i0 + u * NT + threadIdx.x;
                        v[u] = i < P ? (uint32_t)__ldcg(col + i) : 0xffffffffu;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const uint64_t i = i0 + u * NT + threadIdx.x;
                        if (i < P) consider(i, v[u]);
                    }
                }
            }
            __syncthreads();
            hit = s_hit;
        } else {
            for (uint64_t base = 0; base < P; base += NT) {
                const uint64_t i = base + threadIdx.x;
                if (i < P) {
                    const PrfTagPart g = prf_tag_part(T, R, D.tags[i]);
                    const uint32_t ho = prf_low<8, 4>(T, R, g, (uint32_t)chunkId) & cmask;
                    if (ho == (uint32_t)offset) {
                        const uint64_t pp = D.pp[i];
                        if (pp == kDefaultProgramPoint || pp / C != chunkId) atomicMin(&s_hit, (uint32_t)i);
                    }
                }
                __syncthreads();
                hit = s_hit;
                __syncthreads();
                if (hit != 0xffffffffu) break;
            }
            __syncthreads();
        }
        const int status = s_status != 0 ? s_status : (hit == 0xffffffffu ? 4 : 0);   // pir.go:416-419
        if (status != 0) {
            if (threadIdx.x == 0) { meta[t] = ClientMeta{0, 0, status, 0}; a_set[t] = 0; s_hit = 0xffffffffu; }
            __syncthreads();
            continue;
        }
        // ---- phase B: expand the hit hint to a full set (pir.go:424-427), patch it (pir.go:430-439), hand it to the server kernel;
        // the promoted backup hint takes over the slot (pir.go:460-467) ----
        const uint64_t inGroup = s_ingroup, slot = chunkId * M + inGroup, ridx = s_ridx, btag = s_btag;
        if (indexed) {
            const uint64_t pp_hit = mirror ? (uint64_t)s_pp[hit] : (uint64_t)__ldcg(D.pp32 + hit);
            const uint32_t c_pp = pp_hit != kDefaultProgramPoint ? (uint32_t)(pp_hit / C) : 0xffffffffu;
            const uint16_t *row = D.roff + (P + slot) * D.spad;     // offsets of the backup hint, evaluated by the hint kernel
            for (uint32_t c = threadIdx.x; c < S; c += NT) {
                uint32_t v = __ldcg(D.poff + (uint64_t)c * P + hit);
                const uint16_t nv = __ldcg(row + c);
                if (c == c_pp) v = (uint32_t)(pp_hit & cmask);                // pir.go:430-433
                if (c == (uint32_t)chunkId) v = (uint32_t)(ridx & cmask);     // pir.go:436-439
                offsets[(uint64_t)t * stride + c] = v;
                D.poff[(uint64_t)c * P + hit] = nv;
            }
            __syncthreads();   // every thread has read s_pp[hit] before thread 0 replaces it
        } else {
            const PrfTagPart g = prf_tag_part(T, R, D.tags[hit]);
            for (uint32_t c = threadIdx.x; c < S; c += NT) s_offs[c] = prf_low<8, 4>(T, R, g, c) & cmask;
            const uint64_t pp_hit = D.pp[hit];
            __syncthreads();
            if (threadIdx.x == 0) {
                if (pp_hit != kDefaultProgramPoint) s_offs[pp_hit / C] = (uint32_t)(pp_hit & cmask);
                s_offs[chunkId] = (uint32_t)(ridx & cmask);
            }
            __syncthreads();
            for (uint32_t c = threadIdx.x; c < S; c += NT) offsets[(uint64_t)t * stride + c] = s_offs[c];
        }
        if (threadIdx.x == 0) {
            meta[t] = ClientMeta{hit, slot, 0, 0};
            // response-independent half of the refresh (pir.go:460-467)
            D.tags[hit] = btag;
            D.pp[hit] = Q.idx;
            D.pp32[hit] = (uint32_t)Q.idx;
            if (indexed && mirror) s_pp[hit] = (uint32_t)Q.idx;
            s_fin += 1;
            *D.finished = s_fin;
            D.hist[chunkId] = inGroup + 1;
            if (indexed) s_offs[chunkId] = (uint32_t)(inGroup + 1);
            s_hit = 0xffffffffu;
        }
        __threadfence_block();
        __syncthreads();
    }
}


// Thread w owns word w of every entry, so the only cross-query dependency -- two queries of a part refreshing
// the same hint slot -- is a read-after-write inside one thread: no barrier is needed, and the operands that do not
// depend on earlier queries (answer, replacement value, backup parity) are fetched four queries ahead.
// With query vectors given it also evaluates, for every entry it has just finished, L2Dist(vector part of the entry,
// the query vector of that sub-query) -- the per-step L2Dist call site of SearchKNN (search.go:204) without a launch of
// its own: qv = [n_vecs][dim], vid = per-query vector index or nullptr (one vector for all), dist_out[t].
__global__ void __launch_bounds__(256) client_finish_kernel(const ClientPartDev *parts, const uint32_t *part_start,
                                                            const uint32_t *part_items, const ClientMeta *meta, uint32_t E,
                                                            const uint64_t *__restrict__ answers, uint64_t *__restrict__ out,
                                                            const float *__restrict__ qv, const uint32_t *__restrict__ vid, uint32_t dim,
                                                            float *__restrict__ dist_out, const uint32_t *part_map, uint32_t fixed_per,
                                                            uint64_t *const *__restrict__ out_rows) {
    // out_rows (device-built calls): record t's entry goes to out_rows[t] instead of out + t*E -- the search keeps every
    // successful answer in its client's local cache, so the answer is finished straight into its cache slot
    const uint32_t part = part_map ? part_map[blockIdx.x] : blockIdx.x;
    const ClientPartDev &D = parts[part];
    const uint32_t E4 = E & ~3u;  // EntryXor granularity (xorSlices leaves the len%4 tail untouched)
    __shared__ uint32_t s_list[CL_MAX_LIST];
    struct FinMeta { uint32_t hit, slot; int32_t status; };
    __shared__ FinMeta s_meta[CL_MAX_LIST];
    if (!fixed_per && part_start[part] == part_start[part + 1]) return;
    const uint32_t n_mine = client_load_list(part_start, part_items, part, s_list, CL_MAX_LIST, fixed_per);
    for (uint32_t k = threadIdx.x; k < n_mine; k += blockDim.x) {
        const ClientMeta m = meta[s_list[k]];
        s_meta[k] = FinMeta{(uint32_t)m.hit, (uint32_t)m.slot, m.status};
    }
    __syncthreads();
    // the parity of the hint a query consumed is fetched ahead too unless an earlier query of this part refreshed the same
    // hint slot (then it is read at its turn, after that write)
    __shared__ uint8_t s_dep[CL_MAX_LIST];
    for (uint32_t k = threadIdx.x; k < n_mine; k += blockDim.x) {
        bool dep = n_mine > 64;      // quadratic check: small calls only (a search step has <= `per` queries per part)
        for (uint32_t e = 0; e < k && !dep; e++) dep = s_meta[e].status == 0 && s_meta[e].hit == s_meta[k].hit;
        s_dep[k] = dep ? 1 : 0;
    }
    __syncthreads();
    const uint64_t *__restrict__ bparity = D.bparity, *__restrict__ rval = D.rval;
    constexpr int B = 4;
    for (uint32_t w = threadIdx.x; w < E; w += blockDim.x) {
        for (uint32_t k0 = 0; k0 < n_mine; k0 += B) {
            uint64_t a[B], rv[B], bp[B], pp[B];
#pragma unroll
            for (int j = 0; j < B; j++) {
                a[j] = rv[j] = bp[j] = pp[j] = 0;
                if (k0 + j < n_mine && s_meta[k0 + j].status == 0) {
                    const uint64_t slot = s_meta[k0 + j].slot;
                    a[j] = answers[(uint64_t)s_list[k0 + j] * E + w];
                    rv[j] = rval[slot * E + w];
                    bp[j] = bparity[slot * E + w];
                    if (!s_dep[k0 + j] && w < E4) pp[j] = D.parity[(uint64_t)s_meta[k0 + j].hit * E + w];
                }
            }
#pragma unroll
            for (int j = 0; j < B; j++) {
                if (k0 + j >= n_mine) break;
                const uint32_t t = s_list[k0 + j];
                uint64_t *orow = out_rows ? out_rows[t] : out + (uint64_t)t * E;
                if (s_meta[k0 + j].status != 0) {  // dummy or failed: zero entry (pir.go:356-360)
                    orow[w] = 0;
                    continue;
                }
                uint64_t *par = D.parity + (uint64_t)s_meta[k0 + j].hit * E;
                uint64_t r = a[j], np = bp[j];           // copy(primaryParity[hit], backupParity[slot])  pir.go:461
                if (w < E4) {
                    r ^= rv[j] ^ (s_dep[k0 + j] ? par[w] : pp[j]);   // pir.go:451,453
                    np ^= r;                             // pir.go:463
                }
                par[w] = np;
                orow[w] = r;
            }
        }
    }
    if (dist_out == nullptr) return;
    __syncthreads();   // the entries above were written by this CTA: visible to all its threads from here on
    const int half = threadIdx.x & 1;
    const uint32_t n_up = (n_mine + 15) & ~15u;   // keep whole warps in the pair shuffle
    for (uint32_t k = threadIdx.x >> 1; k < n_up; k += blockDim.x >> 1) {
        const bool ok = k < n_mine;
        const uint32_t t = s_list[ok ? k : 0];
        const float *a = reinterpret_cast<const float *>(out_rows ? out_rows[t] : out + (uint64_t)t * E);
        const float *b = qv + (vid ? (uint64_t)vid[t] * dim : 0);
        const float d = l2_pair<false>(a, b, dim, half);
        if (ok && half == 0) dist_out[t] = d;
    }
}

}  // namespace pm

// =============================================================================================
// C-ABI: pm_client_*
// =============================================================================================
struct pm_search;
namespace pm {
int search_clear_cache(pm_search *s, const uint32_t *part_ids, uint64_t n, cudaStream_t st);   // pm_search.cuh
}
struct pm_client {
    pm_search *search = nullptr;   // device-resident search state attached to this client (its local caches follow the tables)
    pm_db *db;
    uint64_t n_parts, E;
    std::vector<pm::ClientPartDev> host_parts;  // device pointers inside
    pm::ClientPartDev *d_parts;
    void *arena;
    uint64_t max_set;
    std::mutex mu;
    cudaStream_t stream = nullptr;          // every client owns its stream, scratch and lock, so several clients that
    void *wbuf[2] = {nullptr, nullptr};     // share one pm_db (one per user) can be driven from different host threads
    size_t wbytes[2] = {0, 0};
    void *stage_in = nullptr;    // pinned host staging for the input block of a call
    size_t stage_in_bytes = 0;
    void *stage = nullptr;   // pinned host staging for results
    size_t stage_bytes = 0;
    cudaEvent_t ev[6] = {};
    double prof_ms[5] = {};
    uint64_t prof_calls = 0;
};

static int client_scratch(pm_client *c, int slot, size_t bytes, void **out) {
    if (c->wbytes[slot] < bytes) {
        if (c->wbuf[slot]) {
            PM_CUDA(cudaStreamSynchronize(c->stream));
            PM_CUDA(cudaFree(c->wbuf[slot]));
            c->wbuf[slot] = nullptr;
            c->wbytes[slot] = 0;
        }
        const size_t want = std::max<size_t>(bytes + bytes / 2, (size_t)1 << 20);   // re-growing costs a cudaFree (device-wide sync)
        cudaError_t e = cudaMalloc(&c->wbuf[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return pm::set_error(PM_ERR_NOMEM, "pm_client: cudaMalloc(%zu) for scratch failed: %s", want, cudaGetErrorString(e));
        }
        c->wbytes[slot] = want;
    }
    *out = c->wbuf[slot];
    return PM_OK;
}

PM_EXPORT int pm_client_create(pm_db *db, const pm_client_part *parts, uint64_t n_parts, pm_client **out) {
    using namespace pm;
    if (!db || !parts || !out || n_parts == 0) return set_error(PM_ERR_ARG, "pm_client_create: null pointer / no parts");
    *out = nullptr;
    int rc = ensure_device(db->device);
    if (rc) return rc;
    const uint64_t E = db->entry_u64;
    uint64_t words = 0, max_set = 0;
    for (uint64_t i = 0; i < n_parts; i++) {
        const pm_client_part &p = parts[i];
        if (p.chunk_size == 0 || (p.chunk_size & (p.chunk_size - 1)) || p.row0 + p.n_rows > db->n_rows || p.set_size == 0 ||
            p.set_size > 16384 || p.chunk_size * p.set_size > 0x7fffffffull || p.n_primary >= 0xffffffffull)
            return set_error(PM_ERR_ARG, "pm_client_create: bad geometry for part %llu", (unsigned long long)i);
        const uint64_t P = p.n_primary, B = p.set_size * p.backup_group;
        words += 2 * P + P * E + 2 * B + 2 * B * E + p.set_size + 2;
        if (p.chunk_size <= 65536) words += (p.set_size * P * 2 + 7) / 8 + 2 + ((P + B) * ((p.set_size + 7) & ~7ull) * 2 + 7) / 8 + 2;   // offset index
        words += (P * 4 + 7) / 8 + 2;                                           // u32 program points
        words = (words + 3) & ~1ull;
        if (p.set_size > max_set) max_set = p.set_size;
    }
    pm_client *c = new (std::nothrow) pm_client();
    if (!c) return set_error(PM_ERR_NOMEM, "out of host memory");
    c->db = db; c->n_parts = n_parts; c->E = E; c->max_set = max_set;
    cudaError_t e = cudaMalloc(&c->arena, words * 8 + 256);
    if (e != cudaSuccess) { cudaGetLastError(); delete c; return set_error(PM_ERR_NOMEM, "pm_client_create: cudaMalloc(%llu) failed", (unsigned long long)(words * 8)); }
    e = cudaMalloc(&c->d_parts, n_parts * sizeof(ClientPartDev));
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(c->arena); delete c; return set_error(PM_ERR_NOMEM, "pm_client_create: cudaMalloc failed"); }
    uint64_t *cur = (uint64_t *)c->arena;
    c->host_parts.resize(n_parts);
    for (uint64_t i = 0; i < n_parts; i++) {
        const pm_client_part &p = parts[i];
        ClientPartDev &D = c->host_parts[i];
        memset(&D, 0, sizeof(D));
        const uint64_t P = p.n_primary, B = p.set_size * p.backup_group;
        D.row0 = p.row0; D.n_rows = p.n_rows; D.chunk_size = p.chunk_size; D.set_size = p.set_size;
        D.n_primary = P; D.backup_group = p.backup_group; D.max_query_num = p.max_query_num;
        D.chunk_mask = (uint32_t)(p.chunk_size - 1); D.chunk_shift = (uint32_t)__builtin_ctzll(p.chunk_size);
        D.parity = cur; cur += P * E;      // primary then backup parities contiguous in hint-number order
        D.bparity = cur; cur += B * E;
        D.rval = cur; cur += B * E;
        D.tags = cur; cur += P;
        D.pp = cur; cur += P;
        D.btags = cur; cur += B;
        D.ridx = cur; cur += B;
        D.hist = cur; cur += p.set_size;
        D.finished = cur; cur += 2;
        if ((uintptr_t)cur & 15) cur += 1;
        D.poff = nullptr;
        D.roff = nullptr;
        D.spad = (uint32_t)((p.set_size + 7) & ~7ull);
        if (p.chunk_size <= 65536) {
            D.poff = (uint16_t *)cur;
            cur += (p.set_size * P * 2 + 7) / 8;
            if ((uintptr_t)cur & 15) cur += 1;
            D.roff = (uint16_t *)cur;
            cur += ((P + B) * D.spad * 2 + 7) / 8;
            if ((uintptr_t)cur & 15) cur += 1;
        }
        D.pp32 = (uint32_t *)cur;
        cur += (P * 4 + 7) / 8;
        if ((uintptr_t)cur & 15) cur += 1;
    }
    // the client's kernels are small and latency bound; they run on a high-priority stream so that, when several clients
    // (or lock-step groups) share the GPU, their CTAs slip in between the CTAs of another group's HBM-bound answer kernel
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
    }
    if (e != cudaSuccess) { cudaFree(c->arena); cudaFree(c->d_parts); delete c; return set_error(PM_ERR_CUDA, "pm_client_create: stream creation failed"); }
    *out = c;
    return PM_OK;
}

PM_EXPORT int pm_client_destroy(pm_client *c) {
    if (!c) return PM_OK;
    if (c->prof_calls)
        fprintf(stderr, "[client profile] %llu calls, us per call: h2d %.1f | prepare %.1f | answer %.1f | finish %.1f | d2h %.1f\n",
                (unsigned long long)c->prof_calls, c->prof_ms[0] / c->prof_calls * 1e3, c->prof_ms[1] / c->prof_calls * 1e3,
                c->prof_ms[2] / c->prof_calls * 1e3, c->prof_ms[3] / c->prof_calls * 1e3, c->prof_ms[4] / c->prof_calls * 1e3);
    if (pm::ensure_device(c->db->device) == PM_OK) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->arena);
        cudaFree(c->d_parts);
        for (int i = 0; i < 2; i++) if (c->wbuf[i]) cudaFree(c->wbuf[i]);
        if (c->stage) cudaFreeHost(c->stage);
        if (c->stage_in) cudaFreeHost(c->stage_in);
        cudaStreamDestroy(c->stream);
    }
    delete c;
    return PM_OK;
}

PM_EXPORT int pm_client_preprocess(pm_client *c, const uint32_t *part_ids, uint64_t n, const uint32_t *rk, const uint64_t *repl_seed,
                                   int skip_prep) {
    using namespace pm;
    if (!c || (n && (!part_ids || !rk || !repl_seed))) return set_error(PM_ERR_ARG, "pm_client_preprocess: null pointer");
    if (n == 0) return PM_OK;
    int rc = ensure_device(c->db->device);
    if (rc) return rc;
    pm_db *db = c->db;
    std::lock_guard<std::mutex> lock(c->mu);
    const uint64_t E = c->E;
    uint64_t max_n = 0;
    for (uint64_t a = 0; a < n; a++) {
        if (part_ids[a] >= c->n_parts) return set_error(PM_ERR_ARG, "pm_client_preprocess: part id out of range");
        ClientPartDev &D = c->host_parts[part_ids[a]];
        memcpy(D.rk, rk + a * 44, sizeof(D.rk));
        const uint64_t P = D.n_primary, B = D.set_size * D.backup_group;
        max_n = std::max(max_n, std::max(P, B));
    }
    PM_CUDA(cudaMemcpyAsync(c->d_parts, c->host_parts.data(), c->n_parts * sizeof(ClientPartDev), cudaMemcpyHostToDevice, c->stream));
    void *d_tmp = nullptr;
    if ((rc = client_scratch(c, 1, n * 16, &d_tmp))) return rc;
    uint32_t *d_ids = (uint32_t *)d_tmp;
    uint64_t *d_seed = (uint64_t *)((char *)d_tmp + ((n * 4 + 7) & ~7ull));
    PM_CUDA(cudaMemcpyAsync(d_ids, part_ids, n * 4, cudaMemcpyHostToDevice, c->stream));
    PM_CUDA(cudaMemcpyAsync(d_seed, repl_seed, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (c->search && (rc = search_clear_cache(c->search, part_ids, n, c->stream))) return rc;   // Initialization clears localCache (pir.go:127)
    dim3 grid((unsigned)std::min<uint64_t>((max_n + 255) / 256, 64), (unsigned)n);
    client_init_kernel<<<grid, 256, 0, c->stream>>>(c->d_parts, d_ids, d_seed, skip_prep);
    PM_CHECK_LAUNCH();
    count_launch();
    if (skip_prep) {   // DummyPreprocessing runs no hint kernel: the offsets of all hints are evaluated here
        uint64_t max_h = 0;
        for (uint64_t a = 0; a < n; a++) {
            const ClientPartDev &D = c->host_parts[part_ids[a]];
            max_h = std::max<uint64_t>(max_h, D.n_primary + D.set_size * D.backup_group);
        }
        dim3 fgrid((unsigned)std::max<uint64_t>(1, (max_h + 255) / 256), (unsigned)n);
        PM_CUDA(cudaFuncSetAttribute(client_fill_roff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, aes_tab_words<1>() * 4));
        client_fill_roff_kernel<<<fgrid, 256, aes_tab_words<1>() * 4, c->stream>>>(c->d_parts, d_ids);
        PM_CHECK_LAUNCH();
        count_launch();
    }
    std::vector<pm_hint_job> jobs;
    for (uint64_t a = 0; a < n; a++) {
        const ClientPartDev &D = c->host_parts[part_ids[a]];
        const uint64_t P = D.n_primary, B = D.set_size * D.backup_group;
        if (skip_prep) {  // DummyPreprocessing (pir.go:520-523): Initialization only, parities and replacement values zero
            PM_CUDA(cudaMemsetAsync(D.parity, 0, (P + 2 * B) * E * 8, c->stream));
            continue;
        }
        pm_hint_job j;
        memset(&j, 0, sizeof(j));
        j.row0 = D.row0; j.n_rows = D.n_rows; j.chunk_size = D.chunk_size; j.set_size = D.set_size;
        memcpy(j.rk, D.rk, sizeof(j.rk));
        j.hint_begin = 0; j.n_hints = P + B; j.n_primary = P; j.backup_group = D.backup_group;
        j.parity_out = D.parity;
        j.offsets_out = D.roff;     // nullptr when chunk_size > 65536 (no 16-bit index)
        jobs.push_back(j);
    }
    if (!jobs.empty()) {
        if ((rc = hintgen_enqueue(db, jobs.data(), jobs.size(), c->stream))) return rc;
        for (uint64_t a = 0; a < n; a++) {  // replacement values (pir.go:348)
            const ClientPartDev &D = c->host_parts[part_ids[a]];
            if ((rc = gather_enqueue(db, D.row0, D.n_rows, D.ridx, D.set_size * D.backup_group, D.rval, c->stream))) return rc;
        }
    }
    {   // primary rows of roff -> the column-major index the hint search scans
        uint64_t max_p = 0;
        for (uint64_t a = 0; a < n; a++) max_p = std::max<uint64_t>(max_p, c->host_parts[part_ids[a]].n_primary);
        dim3 tgrid((unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (max_p + 31) / 32), 256), 8, (unsigned)n);
        client_transpose_kernel<<<tgrid, 256, 0, c->stream>>>(c->d_parts, d_ids);
        PM_CHECK_LAUNCH();
        count_launch();
    }
    PM_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}

static int client_query_impl(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                             const float *query_vec, uint64_t n_vecs, const uint32_t *vec_id, uint64_t dim, float *dist_out);
PM_EXPORT int pm_client_query_batch(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status) {
    return client_query_impl(c, queries, q, out, status, nullptr, 0, nullptr, 0, nullptr);
}
PM_EXPORT int pm_client_query_batch_l2(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                                       const float *query_vec, uint64_t dim, float *dist_out) {
    if (!query_vec || !dist_out || dim == 0) return pm::set_error(PM_ERR_ARG, "pm_client_query_batch_l2: null pointer");
    if (c && dim * 4 > c->E * 8) return pm::set_error(PM_ERR_ARG, "pm_client_query_batch_l2: dim does not fit in an entry");
    return client_query_impl(c, queries, q, out, status, query_vec, 1, nullptr, dim, dist_out);
}
PM_EXPORT int pm_client_query_batch_l2m(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                                        const float *query_vecs, uint64_t n_vecs, const uint32_t *vec_id, uint64_t dim, float *dist_out) {
    if (!query_vecs || !vec_id || !dist_out || dim == 0 || n_vecs == 0) return pm::set_error(PM_ERR_ARG, "pm_client_query_batch_l2m: null pointer");
    if (c && dim * 4 > c->E * 8) return pm::set_error(PM_ERR_ARG, "pm_client_query_batch_l2m: dim does not fit in an entry");
    for (uint64_t t = 0; t < q; t++)
        if (vec_id[t] >= n_vecs) return pm::set_error(PM_ERR_ARG, "pm_client_query_batch_l2m: vec_id out of range");
    return client_query_impl(c, queries, q, out, status, query_vecs, n_vecs, vec_id, dim, dist_out);
}
static int client_query_impl(pm_client *c, const pm_client_query *queries, uint64_t q, uint64_t *out, int32_t *status,
                             const float *query_vec, uint64_t n_vecs, const uint32_t *vec_id, uint64_t dim, float *dist_out) {
    using namespace pm;
    if (!c || (q && (!queries || !out || !status))) return set_error(PM_ERR_ARG, "pm_client_query_batch: null pointer");
    if (q == 0) return PM_OK;
    if (q > 1u << 20) return set_error(PM_ERR_UNSUPPORTED, "pm_client_query_batch: too many queries in one call");
    static_assert(sizeof(pm_client_query) == sizeof(ClientQueryDev), "pm_client_query layout");
    std::vector<uint32_t> per_part(c->n_parts, 0);
    for (uint64_t t = 0; t < q; t++) {
        if (queries[t].part >= c->n_parts) return set_error(PM_ERR_ARG, "pm_client_query_batch: part id out of range");
        if (++per_part[queries[t].part] > CL_MAX_LIST)
            return set_error(PM_ERR_UNSUPPORTED, "pm_client_query_batch: more than %u queries for one sub-PIR in one call", CL_MAX_LIST);
        if (queries[t].kind == 1 && queries[t].idx >= c->host_parts[queries[t].part].n_rows)
            return set_error(PM_ERR_ARG, "pm_client_query_batch: idx %llu is out of range", (unsigned long long)queries[t].idx);  // pir.go:373-378
    }
    int rc = ensure_device(c->db->device);
    if (rc) return rc;
    pm_db *db = c->db;
    std::lock_guard<std::mutex> lock(c->mu);
    const uint64_t E = c->E, stride = (c->max_set + 3) & ~3ull;
    // device input block, filled by ONE host-to-device copy from page-locked staging:  queries | per-part lists | query
    // vectors | vector ids;  behind it the offsets and answer descriptors the prepare kernel produces.  Second buffer:
    // answers | results | meta | distances (results, meta and distances leave in one copy).
    const size_t b_q = q * sizeof(ClientQueryDev), b_meta = q * sizeof(ClientMeta), b_off = q * stride * 4, b_desc = q * 24;
    void *d_in = nullptr, *d_out = nullptr;
    const size_t b_qv = (n_vecs * dim * 4 + 15) & ~15ull, b_vid = vec_id ? (q * 4 + 15) & ~15ull : 0, b_dist = dist_out ? q * 4 : 0;
    const size_t b_csr = ((c->n_parts + 1 + q) * 4 + 15) & ~15ull;
    const size_t b_hdr = b_q + b_csr + b_qv + b_vid;   // sizeof(ClientQueryDev) = 32: every piece stays 16-byte aligned
    if ((rc = client_scratch(c, 1, b_hdr + b_off + b_desc + 64, &d_in))) return rc;
    if ((rc = client_scratch(c, 0, 2 * q * E * 8 + b_meta + b_dist, &d_out))) return rc;
    ClientQueryDev *d_q = (ClientQueryDev *)d_in;
    uint32_t *d_start = (uint32_t *)((char *)d_in + b_q), *d_items = d_start + c->n_parts + 1;
    float *d_qv = (float *)((char *)d_in + b_q + b_csr);
    uint32_t *d_vid = vec_id ? (uint32_t *)((char *)d_qv + b_qv) : nullptr;
    uint32_t *d_off = (uint32_t *)((char *)d_in + b_hdr);
    uint64_t *d_row0 = (uint64_t *)((char *)d_off + b_off), *d_nrows = d_row0 + q;
    uint32_t *d_chunk = (uint32_t *)(d_nrows + q), *d_set = d_chunk + q;
    uint64_t *d_ans = (uint64_t *)d_out, *d_res = d_ans + q * E;
    ClientMeta *d_meta = (ClientMeta *)(d_res + q * E);
    float *d_dist = (float *)(d_meta + q);
    if (c->stage_in_bytes < b_hdr) {
        if (c->stage_in) cudaFreeHost(c->stage_in);
        c->stage_in = nullptr;
        c->stage_in_bytes = 0;
        PM_CUDA(cudaHostAlloc(&c->stage_in, b_hdr * 2 + 4096, cudaHostAllocDefault));
        c->stage_in_bytes = b_hdr * 2 + 4096;
    }
    // a page-locked `out` (pm_host_alloc) receives the answers straight from the GPU; only meta + distances are staged
    bool direct = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess) direct = at.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    const size_t b_back = (direct ? 0 : q * E * 8) + b_meta + b_dist;
    if (c->stage_bytes < b_back) {
        if (c->stage) cudaFreeHost(c->stage);
        c->stage = nullptr;
        c->stage_bytes = 0;
        PM_CUDA(cudaHostAlloc(&c->stage, b_back + b_back / 2, cudaHostAllocDefault));
        c->stage_bytes = b_back + b_back / 2;
    }
    // PM_CLIENT_PROFILE=1: CUDA-event breakdown of the call (H2D | prepare | answer | finish | D2H), printed at destroy
    static const bool prof = getenv("PM_CLIENT_PROFILE") != nullptr;
    if (prof && !c->ev[0]) for (int i = 0; i < 6; i++) cudaEventCreate(&c->ev[i]);
    auto mark = [&](int i) { if (prof) cudaEventRecord(c->ev[i], c->stream); };
    // fill the input block: records, per-part query lists (counting sort; array order within a part is the processing
    // order), query vectors and their ids
    {
        char *h = (char *)c->stage_in;
        memcpy(h, queries, b_q);
        uint32_t *start = (uint32_t *)(h + b_q), *items = start + c->n_parts + 1;
        uint32_t acc = 0;
        for (uint64_t i = 0; i < c->n_parts; i++) { start[i] = acc; acc += per_part[i]; }
        start[c->n_parts] = acc;
        std::vector<uint32_t> &cur = per_part;   // reuse as running positions
        for (uint64_t i = 0; i < c->n_parts; i++) cur[i] = start[i];
        for (uint64_t t = 0; t < q; t++) items[cur[queries[t].part]++] = (uint32_t)t;
        for (uint64_t i = 0; i < c->n_parts; i++) cur[i] -= start[i];   // back to counts
        if (dist_out) memcpy(h + b_q + b_csr, query_vec, n_vecs * dim * 4);
        if (vec_id) memcpy(h + b_q + b_csr + b_qv, vec_id, q * 4);
    }
    mark(0);
    PM_CUDA(cudaMemcpyAsync(d_in, c->stage_in, b_hdr, cudaMemcpyHostToDevice, c->stream));
    mark(1);
    uint64_t max_p = 0;
    for (uint64_t i = 0; i < c->n_parts; i++) if (c->host_parts[i].poff) max_p = std::max<uint64_t>(max_p, c->host_parts[i].n_primary);
    // the program points are mirrored in shared memory when they fit; larger instances (P = 59392 for 2^20 rows) read the
    // u32 copy in global memory on the few offset matches of a column scan instead
    const size_t smem_base = (aes_tab_words<8>() + 64 + stride + CL_MAX_LIST) * 4;
    const uint32_t mirror = smem_base + max_p * 4 <= 200 * 1024 ? 1u : 0u;
    const size_t smem = smem_base + (mirror ? max_p * 4 : 0);
    if (smem > 200 * 1024) return set_error(PM_ERR_UNSUPPORTED, "pm_client_query_batch: set_size too large for the prepare kernel's shared memory");
    PM_CUDA(cudaFuncSetAttribute(client_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 512-thread CTAs (2 per SM by registers) give the shortest call for one client (16 parts: 37 vs 48 us); clients
    // that carry a lock-step group run 256-thread CTAs (4 per SM): more parts -- of this call and of the other groups'
    // concurrent calls -- are resident at once (measured: one 32-lane group 139 -> 116 us, 4 x 16 lanes +5 % queries/s)
    const unsigned prep_threads = c->n_parts <= 64 ? CL_THREADS : CL_THREADS / 2;
    client_prepare_kernel<<<(unsigned)c->n_parts, prep_threads, smem, c->stream>>>(c->d_parts, d_q, d_start, d_items, (uint32_t)q, (uint32_t)stride,
                                                                                  d_off, d_meta, d_row0, d_nrows, d_chunk, d_set, mirror, nullptr, 0);
    PM_CHECK_LAUNCH();
    count_launch();
    mark(2);
    if ((rc = answer_enqueue(db, d_row0, d_nrows, d_chunk, d_set, d_off, stride, q, (uint32_t)stride, d_ans, c->stream))) return rc;
    mark(3);
    // the distances of the answered entries' vectors to the search query (A10 call site) come out of the same launch
    client_finish_kernel<<<(unsigned)c->n_parts, 256, 0, c->stream>>>(c->d_parts, d_start, d_items, d_meta, (uint32_t)E, d_ans, d_res,
                                                                      dist_out ? d_qv : nullptr, d_vid, (uint32_t)dim, dist_out ? d_dist : nullptr, nullptr, 0, nullptr);
    PM_CHECK_LAUNCH();
    count_launch();
    mark(4);
    if (direct) {
        PM_CUDA(cudaMemcpyAsync(out, d_res, q * E * 8, cudaMemcpyDeviceToHost, c->stream));
        PM_CUDA(cudaMemcpyAsync(c->stage, d_meta, b_back, cudaMemcpyDeviceToHost, c->stream));
    } else {
        PM_CUDA(cudaMemcpyAsync(c->stage, d_res, b_back, cudaMemcpyDeviceToHost, c->stream));
    }
    mark(5);
    PM_CUDA(cudaStreamSynchronize(c->stream));
    const size_t res_in_stage = direct ? 0 : q * E * 8;
    if (!direct) memcpy(out, c->stage, q * E * 8);
    const ClientMeta *meta = (const ClientMeta *)((const char *)c->stage + res_in_stage);
    if (dist_out) memcpy(dist_out, (const char *)c->stage + res_in_stage + b_meta, q * 4);
    if (prof) {
        for (int i = 0; i < 5; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]);
            c->prof_ms[i] += ms;
        }
        c->prof_calls++;
    }
    for (uint64_t t = 0; t < q; t++) status[t] = meta[t].status < 0 ? 0 : meta[t].status;
    return PM_OK;
}

PM_EXPORT int pm_host_alloc(void **out, uint64_t bytes) {
    if (!out) return pm::set_error(PM_ERR_ARG, "pm_host_alloc: null pointer");
    *out = nullptr;
    if (bytes == 0) return PM_OK;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return pm::set_error(PM_ERR_NOMEM, "pm_host_alloc: cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    return PM_OK;
}
PM_EXPORT int pm_host_free(void *p) {
    if (p) PM_CUDA(cudaFreeHost(p));
    return PM_OK;
}

PM_EXPORT int pm_client_download(pm_client *c, uint32_t part, int table, uint64_t *out, uint64_t cap_words) {
    using namespace pm;
    if (!c || !out || part >= c->n_parts) return set_error(PM_ERR_ARG, "pm_client_download: bad argument");
    int rc = ensure_device(c->db->device);
    if (rc) return rc;
    const ClientPartDev &D = c->host_parts[part];
    const uint64_t P = D.n_primary, B = D.set_size * D.backup_group, E = c->E;
    const uint64_t *src = nullptr;
    uint64_t words = 0;
    switch (table) {
    case 0: src = D.tags; words = P; break;
    case 1: src = D.parity; words = P * E; break;
    case 2: src = D.pp; words = P; break;
    case 3: src = D.ridx; words = B; break;
    case 4: src = D.rval; words = B * E; break;
    case 5: src = D.btags; words = B; break;
    case 6: src = D.bparity; words = B * E; break;
    case 7: src = D.hist; words = D.set_size; break;
    case 8: src = D.finished; words = 1; break;
    default: return set_error(PM_ERR_ARG, "pm_client_download: unknown table %d", table);
    }
    if (cap_words < words) return set_error(PM_ERR_ARG, "pm_client_download: buffer too small (%llu < %llu words)",
                                            (unsigned long long)cap_words, (unsigned long long)words);
    std::lock_guard<std::mutex> lock(c->mu);
    PM_CUDA(cudaMemcpyAsync(out, src, words * 8, cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}
