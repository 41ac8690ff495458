// GPU-resident lock-step SearchKNN: the frontier of graphann.SearchKNN (search.go:114-234) --
// explore heap, known set, reach steps, final re-rank -- and the bookkeeping of SimpleBatchPianoPIR.Query around the
// fetch (batch-pir.go:170-248: bucketing by partition, surplus drops, dummy padding, the client's local cache,
// pir.go:381-383/468) live in HBM, one lane per independent client of a pm_client group.  A search step is four launches
// on the client's stream and the host only enqueues them:
//     search_step_kernel     apply the previous fetch (cache put, fresh vertices, distances of cache hits, heap pushes),
//                            then pop `parallel` vertices, emit their neighbour lists as the next batch and build the
//                            pm_client_query records of every lane straight from the neighbour lists in HBM
//     client_prepare_kernel  |
//     answer_kernel          |  pm_client.cuh / pm_pir.cu, unchanged arithmetic (fixed record layout, part map)
//     client_finish_kernel   |  (+ the distance of every answered vector to its lane's query)
// Entries never travel to the host; a round of searches returns k ids and reach steps per lane.
// Tie rules are those of the host mirror and the oracle: start ranking by (distance, position), explore queue
// = container/heap's sift-up / sift-down on `dist <`, final ranking by (distance, id).  Every lane returns exactly what
// its client returns when it searches alone (tests/test_search_device_gpu.py compare with the CPU oracle per lane).
#pragma once

namespace pm {

struct SearchDev {
    // geometry
    uint32_t n, dim, m, E, PN, per, B, R, ns, cap, hcap, sort_cap, cache_cap, cache_hcap, max_step, parallel, recq_off;
    uint64_t PS;
    const uint64_t *db;              // packed entries: dim f32 || m u32 per row
    const ClientPartDev *parts;
    // per lane [L][...]
    uint64_t *lane_u64;              // [L][8]: rseed, rctr, total_q, succ_q, server_q
    uint32_t *lane_u32;              // [L][8]: known_cnt, heap_cnt, step
    uint32_t *known_id, *known_step, *hash_key, *hash_val, *heap_slot, *nbr, *batch_ids;
    int32_t *batch_rec;
    float *known_dist, *heap_dist, *qvec;
    uint32_t *start_id, *start_nbr;
    float *start_vec, *start_dist;
    // per part [L*PN]
    uint32_t *cache_key, *cache_val, *cache_cnt;
    uint64_t *cache_entries, *dummy_seed, *dummy_ctr;
    // per call (indexed by the lane's position `a` in the round)
    ClientQueryDev *records;         // [act][R]
    const ClientMeta *meta;
    uint64_t *results;               // [act][R][E] scratch rows (dummy / skipped records)
    uint64_t **out_rows;             // [act][R] where the finish kernel puts record r's entry: its cache slot (real queries) or a scratch row
    const float *dist;               // [act][R]
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t x, uint32_t mask) { return (x * 2654435761u >> 7) & mask; }
__device__ __forceinline__ int hash_find(const uint32_t *keys, const uint32_t *vals, uint32_t mask, uint32_t key) {
    for (uint32_t h = hash_u32(key, mask);; h = (h + 1) & mask) {
        const uint32_t k = keys[h];
        if (k == key) return (int)vals[h];
        if (k == 0xffffffffu) return -1;
    }
}
// single-writer insert (the caller serialises writers of one table)
__device__ __forceinline__ void hash_put(uint32_t *keys, uint32_t *vals, uint32_t mask, uint32_t key, uint32_t val) {
    for (uint32_t h = hash_u32(key, mask);; h = (h + 1) & mask) {
        const uint32_t k = keys[h];
        if (k == key || k == 0xffffffffu) { keys[h] = key; vals[h] = val; return; }
    }
}

// container/heap (search.go:92-111): Push = append + up, Pop = swap(0, n-1) + down(0, n-1) + remove last; Less = dist <
struct HeapView { float *d; uint32_t *s; uint32_t n; };
// Both walks keep the moving element in registers and shift the others (the arrangement they leave is the one the
// reference's swap-based up/down leaves: the same comparisons in the same order decide where the element stops).
__device__ void heap_push(HeapView &h, float dist, uint32_t slot) {
    uint32_t j = h.n++;
    while (j > 0) {
        const uint32_t i = (j - 1) / 2;
        const float di = h.d[i];
        if (!(dist < di)) break;
        h.d[j] = di; h.s[j] = h.s[i];
        j = i;
    }
    h.d[j] = dist; h.s[j] = slot;
}
// The same push by one whole warp (every lane calls it with the same arguments; n = entries before the push, the caller
// counts).  The walk of heap_push only READS the ancestors of position n -- what it writes lies below what it still has to
// read -- so the chain can be read at once: lane l holds the ancestor l+1 levels up, the first lane whose ancestor is not
// greater than the new element is where the sequential walk stops, the ancestors below it move down one level each and the
// new element takes the freed position.  Same comparisons, same final arrangement, one step instead of up to log2(n).
// Pushes follow each other through shared memory, so the push-to-push chain is kept short: 32-bit shared addresses
// (hd, hs = __cvta_generic_to_shared of the two arrays), lane masks instead of a find-first-set, predicated stores.
__device__ __forceinline__ void heap_push_warp(uint32_t hd, uint32_t hs, uint32_t n, float dist, uint32_t slot, uint32_t lanemask_lt, uint32_t lanemask_le) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mine = (n + 1) >> lane;        // 1-based position `lane` levels above the new leaf (0 = above the root); n + 1 < 2^31
    const uint32_t up = mine >> 1;                // its parent, 0 = none
    const uint32_t ua = (up ? up - 1 : 0) * 4;
    float du;
    uint32_t su;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(du) : "r"(hd + ua) : "memory");
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(su) : "r"(hs + ua) : "memory");
    const uint32_t stop = __ballot_sync(0xffffffffu, !up || !(dist < du));       // the lane at the root always stops
    const uint32_t wr = (stop & lanemask_lt) == 0;                                // lane <= first stopping lane: this position changes
    const bool shift = (stop & lanemask_le) == 0;                                 // lane <  first stopping lane: its parent moves down into it
    const float vd = shift ? du : dist;
    const uint32_t vs = shift ? su : slot;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p st.shared.f32 [%0], %1;\n\t@p st.shared.u32 [%2], %3;\n\t}"
                 ::"r"(hd + (mine - 1) * 4), "f"(vd), "r"(hs + (mine - 1) * 4), "r"(vs), "r"(wr) : "memory");
    __syncwarp();
}
__device__ uint32_t heap_pop(HeapView &h) {
    const uint32_t n = h.n - 1, top = h.s[0];
    const float x = h.d[n];          // heap.Pop: Swap(0, n-1), down(0, n-1): the last element sinks from the root
    const uint32_t xs = h.s[n];
    uint32_t i = 0;
    for (;;) {
        const uint32_t j1 = 2 * i + 1;
        if (j1 >= n) break;
        uint32_t j = j1;
        float dj = h.d[j1];
        if (j1 + 1 < n) {
            const float d2 = h.d[j1 + 1];
            if (d2 < dj) { j = j1 + 1; dj = d2; }
        }
        if (!(dj < x)) break;
        h.d[i] = dj; h.s[i] = h.s[j];
        i = j;
    }
    if (n > 0) { h.d[i] = x; h.s[i] = xs; }
    h.n = n;
    return top;
}

constexpr int SR_THREADS = 128;

// distances of every start vertex of a lane to the lane's query (search.go:131-134)
__global__ void __launch_bounds__(SR_THREADS) search_start_dist_kernel(SearchDev S, const uint32_t *lane_ids) {
    const uint32_t lane = lane_ids[blockIdx.x];
    const float *q = S.qvec + (uint64_t)blockIdx.x * S.dim;    // query vectors are indexed by the lane's position in the round
    const int half = threadIdx.x & 1;
    const uint32_t ns_up = (S.ns + 15) & ~15u;
    for (uint32_t i = blockIdx.y * (SR_THREADS / 2) + (threadIdx.x >> 1); i < ns_up; i += gridDim.y * (SR_THREADS / 2)) {
        const bool ok = i < S.ns;
        const float d = l2_pair<false>(S.start_vec + ((uint64_t)lane * S.ns + (ok ? i : 0)) * S.dim, q, S.dim, half);
        if (ok && half == 0) S.start_dist[(uint64_t)lane * S.ns + i] = d;
    }
}

// search.go:117-148: reset the lane's search state, rank the start vertices by (distance, position) and take the first
// `parallel` distinct ones as known vertices and explore-queue entries.  One CTA per lane.
__global__ void __launch_bounds__(SR_THREADS) search_begin_kernel(SearchDev S, const uint32_t *lane_ids, const uint64_t *rseeds, uint32_t benchmarking) {
    extern __shared__ unsigned long long s_key[];   // [ns] (distance bits << 32) | position
    __shared__ unsigned long long s_red[SR_THREADS / 32];
    __shared__ unsigned long long s_best;
    const uint32_t lane = lane_ids[blockIdx.x], t = threadIdx.x;
    uint32_t *hk = S.hash_key + (uint64_t)lane * S.hcap, *hv = S.hash_val + (uint64_t)lane * S.hcap;
    for (uint32_t i = t; i < S.hcap; i += SR_THREADS) hk[i] = 0xffffffffu;
    uint64_t *l64 = S.lane_u64 + (uint64_t)lane * 8;
    uint32_t *l32 = S.lane_u32 + (uint64_t)lane * 8;
    if (t == 0) { l64[0] = rseeds[blockIdx.x]; l64[1] = 0; l64[2] = 0; l64[3] = 0; l64[4] = 0; l32[2] = 0; }
    if (benchmarking) {
        if (t == 0) { l32[0] = 0; l32[1] = 0; }
        return;
    }
    const float *sd = S.start_dist + (uint64_t)lane * S.ns;
    for (uint32_t i = t; i < S.ns; i += SR_THREADS) s_key[i] = ((unsigned long long)__float_as_uint(sd[i]) << 32) | i;   // squared L2 >= +0: bit order = value order
    __syncthreads();
    HeapView hp{S.heap_dist + (uint64_t)lane * S.cap, S.heap_slot + (uint64_t)lane * S.cap, 0};
    uint32_t known = 0;
    for (uint32_t it = 0; it < S.ns && hp.n < S.parallel; it++) {
        unsigned long long best = ~0ull;
        for (uint32_t i = t; i < S.ns; i += SR_THREADS) best = min(best, s_key[i]);
        for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((t & 31) == 0) s_red[t >> 5] = best;
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < SR_THREADS / 32; w++) best = min(best, s_red[w]);
            s_best = best;
        }
        __syncthreads();
        best = s_best;
        if (best == ~0ull) break;
        const uint32_t pos = (uint32_t)best;
        const uint32_t id = S.start_id[(uint64_t)lane * S.ns + pos];
        const bool dup = hash_find(hk, hv, S.hcap - 1, id) >= 0;   // every thread reads the same table: uniform
        __syncthreads();
        if (t == 0) s_key[pos] = ~0ull;
        if (!dup) {
            const float d = __uint_as_float((uint32_t)(best >> 32));
            if (t == 0) {
                hash_put(hk, hv, S.hcap - 1, id, known);
                S.known_id[(uint64_t)lane * S.cap + known] = id;
                S.known_dist[(uint64_t)lane * S.cap + known] = d;
                S.known_step[(uint64_t)lane * S.cap + known] = 0;
                heap_push(hp, d, known);
            }
            for (uint32_t j = t; j < S.m; j += SR_THREADS)
                S.nbr[((uint64_t)lane * S.cap + known) * S.m + j] = S.start_nbr[((uint64_t)lane * S.ns + pos) * S.m + j];
            known++;
            hp.n = known;   // threads other than 0 track the count only (one push per accepted vertex)
        }
        __syncthreads();
    }
    if (t == 0) { l32[0] = known; l32[1] = known; }
}

// One CTA per lane: [apply the previous fetch] then [emit the next batch and its records].  See the file header.
__global__ void __launch_bounds__(SR_THREADS) search_step_kernel(SearchDev S, const uint32_t *lane_ids, uint32_t apply, uint32_t next, uint32_t benchmarking) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t lane = lane_ids[blockIdx.x], a = blockIdx.x, t = threadIdx.x;
    const uint32_t B = S.B, R = S.R, m = S.m, E = S.E;
    // shared layout: heap dist [cap] | heap slot [cap] | ids, aux, flags, fdist, pushd, pushs [B] each | record info 6 x [R] | src [parallel] | cnt [PN]
    float *s_hd = reinterpret_cast<float *>(s_raw);
    uint32_t *s_hs = reinterpret_cast<uint32_t *>(s_hd + S.cap);
    uint32_t *s_ids = s_hs + S.cap;
    uint32_t *s_aux = s_ids + B;       // apply: entry locator / rank ; next: partition rank
    uint32_t *s_flag = s_aux + B;
    float *s_fd = reinterpret_cast<float *>(s_flag + B);
    float *s_pushd = s_fd + B;         // (distance, slot) of the fresh vertices in batch order
    uint32_t *s_pushs = reinterpret_cast<uint32_t *>(s_pushd + B);
    uint32_t *s_rkind = s_pushs + B, *s_rpart = s_rkind + R, *s_ridx = s_rpart + R, *s_rref = s_ridx + R, *s_rok = s_rref + R, *s_rslot = s_rok + R;
    uint32_t *s_src = s_rslot + R;
    uint32_t *s_cnt = s_src + S.parallel;
    ClientQueryDev *s_recq = reinterpret_cast<ClientQueryDev *>(s_raw + S.recq_off);   // records being built
    __shared__ uint32_t s_known, s_nfresh, s_succ, s_server, s_heap_n;
    uint64_t *l64 = S.lane_u64 + (uint64_t)lane * 8;
    uint32_t *l32 = S.lane_u32 + (uint64_t)lane * 8;
    float *g_hd = S.heap_dist + (uint64_t)lane * S.cap;
    uint32_t *g_hs = S.heap_slot + (uint64_t)lane * S.cap;
    uint32_t *hk = S.hash_key + (uint64_t)lane * S.hcap, *hv = S.hash_val + (uint64_t)lane * S.hcap;
    uint32_t heap_n = l32[1];
    // cap is a multiple of 4 and every array 16-byte aligned: the heap moves as uint4 (the words past heap_n are don't-cares)
    for (uint32_t i = t * 4; i < heap_n; i += SR_THREADS * 4) {
        *reinterpret_cast<uint4 *>(s_hd + i) = *reinterpret_cast<const uint4 *>(g_hd + i);
        *reinterpret_cast<uint4 *>(s_hs + i) = *reinterpret_cast<const uint4 *>(g_hs + i);
    }
    if (t == 0) { s_known = l32[0]; s_nfresh = 0; s_succ = 0; s_server = 0; s_heap_n = heap_n; }
    __syncthreads();
    const ClientQueryDev *recs = S.records + (uint64_t)a * R;

    if (apply) {
        const ClientMeta *meta = S.meta + (uint64_t)a * R;
        // the records of the fetch being applied, once into shared memory
        for (uint32_t r = t; r < R; r += SR_THREADS) {
            const ClientQueryDev q = recs[r];
            s_rkind[r] = q.kind;
            s_rpart[r] = q.part;
            s_ridx[r] = (uint32_t)q.idx;
            s_rref[r] = (uint32_t)q.dummy_seed;
            s_rok[r] = (q.kind == 1 && meta[r].status == 0) ? 1u : 0u;
        }
        for (uint32_t i = t; i < B; i += SR_THREADS) s_ids[i] = S.batch_ids[(uint64_t)lane * B + i];
        __syncthreads();
        // (1) settle (batch-pir.go:189-216 / pir.go:468): every successful real answer enters its sub-PIR's local cache.  The
        // finish kernel has already written it into its cache slot (chosen when the record was built): only the key is
        // published here (atomicCAS claims; the keys of one call are distinct).  A failed query leaves its slot unused.
        for (uint32_t r = t; r < R; r += SR_THREADS) {
            if (s_rkind[r] == 0 || s_rok[r]) atomicAdd(&s_server, 1u);
            if (!s_rok[r]) continue;
            const uint32_t gp = s_rpart[r];
            uint32_t *ck = S.cache_key + (uint64_t)gp * S.cache_hcap, *cv = S.cache_val + (uint64_t)gp * S.cache_hcap;
            const uint32_t key = s_ridx[r], mask = S.cache_hcap - 1;
            for (uint32_t h = hash_u32(key, mask);; h = (h + 1) & mask) {
                const uint32_t old = atomicCAS(&ck[h], 0xffffffffu, key);
                if (old == 0xffffffffu || old == key) { cv[h] = s_rref[r]; break; }
            }
        }
        // (2) the response of every input (batch-pir.go:218-237: keyed by global index, zero rows for drops and failures)
        const bool vec4 = (S.dim % 4 == 0) && (m % 4 == 0) && (E % 2 == 0);
        for (uint32_t i = t; i < B; i += SR_THREADS) {
            const uint32_t id = s_ids[i];
            int r = S.batch_rec[(uint64_t)lane * B + i];
            const uint64_t *e = nullptr;     // nullptr = zero entry
            uint32_t loc = 0xffffffffu;      // where the entry lies: record index; bit 31 = in the cache slot that record names
            bool have_dist = false;
            float d = 0.f;
            if (r >= 0) {
                if (s_rkind[r] == 3) r = (int)s_rref[r];      // same index earlier in this call: that record's answer
                if (s_rkind[r] == 1) {
                    if (s_rok[r]) {
                        e = S.cache_entries + ((uint64_t)s_rpart[r] * S.cache_cap + s_rref[r]) * E;   // finished in place
                        loc = (uint32_t)r;
                        if (!benchmarking) { d = S.dist[(uint64_t)a * R + r]; have_dist = true; }   // no distances are computed in benchmark mode
                    }
                } else if (s_rkind[r] == 2) {
                    e = S.cache_entries + ((uint64_t)s_rpart[r] * S.cache_cap + s_rref[r]) * E;
                    loc = 0x80000000u | (uint32_t)r;      // bit 31: no distance came with it
                }
            }
            // the reference's correctness accounting (private-search.go:480-504) and the failed-fetch test (search.go:192-199)
            const uint32_t *want = reinterpret_cast<const uint32_t *>(S.db + (uint64_t)id * E) + S.dim;
            bool correct = true, ok = false;
            if (e == nullptr) {
                for (uint32_t j = 0; j < m; j++) correct = correct && want[j] == 0;
            } else {
                const uint32_t *nb = reinterpret_cast<const uint32_t *>(e) + S.dim;
                if (vec4) {
                    for (uint32_t j = 0; j < m; j += 4) {
                        const uint4 x = *reinterpret_cast<const uint4 *>(nb + j), y = *reinterpret_cast<const uint4 *>(want + j);
                        correct = correct && x.x == y.x && x.y == y.y && x.z == y.z && x.w == y.w;
                        ok = ok || (x.x | x.y | x.z | x.w) != 0;
                    }
                } else {
                    for (uint32_t j = 0; j < m; j++) {
                        correct = correct && nb[j] == want[j];
                        ok = ok || nb[j] != 0;
                    }
                }
            }
            if (correct) atomicAdd(&s_succ, 1u);
            bool fresh = !benchmarking && ok && hash_find(hk, hv, S.hcap - 1, id) < 0;
            if (fresh) {        // a repeated id is known by its second occurrence (no early exit: the loads stay independent)
                bool rep = false;
#pragma unroll 8
                for (uint32_t j = 0; j < i; j++) rep |= s_ids[j] == id;
                fresh = !rep;
            }
            s_flag[i] = fresh ? (have_dist ? 1u : 2u) : 0u;
            s_fd[i] = d;
            s_aux[i] = loc;
        }
        __syncthreads();
        // (3) distances the fetch did not provide (entries served from the local cache): L2Dist to the lane's query
        {
            const float *q = S.qvec + (uint64_t)a * S.dim;
            const int half = t & 1;
            const uint32_t B_up = (B + 15) & ~15u;
            for (uint32_t i0 = 0; i0 < B_up; i0 += SR_THREADS / 2) {
                const uint32_t i = i0 + (t >> 1);
                const bool need = i < B && s_flag[i] == 2u;
                if (!__any_sync(0xffffffffu, need)) continue;
                const uint64_t *e = S.results;   // any readable row for the pairs of this warp that have nothing to do
                if (need) {
                    const uint32_t r = s_aux[i] & 0x7fffffffu;
                    e = S.cache_entries + ((uint64_t)s_rpart[r] * S.cache_cap + s_rref[r]) * E;
                }
                const float d = l2_pair<false>(reinterpret_cast<const float *>(e), q, S.dim, half);
                if (need && half == 0) s_fd[i] = d;
            }
        }
        __syncthreads();
        // (4) fresh vertices become known, in batch order (search.go:200-206): slot = known count + rank among the fresh
        const uint32_t known0 = s_known, this_step = l32[2];
        for (uint32_t i = t; i < B; i += SR_THREADS) {
            if (!s_flag[i]) continue;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < i; j++) rank += s_flag[j] ? 1u : 0u;
            const uint32_t slot = known0 + rank;
            s_pushd[rank] = s_fd[i];
            s_pushs[rank] = slot;
            s_rslot[rank] = s_aux[i] & 0x7fffffffu;      // the record whose cache slot holds this vertex's entry
            atomicAdd(&s_nfresh, 1u);
            if (slot >= S.cap) continue;     // cannot happen: cap = B*max_step + parallel
            S.known_id[(uint64_t)lane * S.cap + slot] = s_ids[i];
            S.known_dist[(uint64_t)lane * S.cap + slot] = s_fd[i];
            S.known_step[(uint64_t)lane * S.cap + slot] = this_step;
            const uint32_t key = s_ids[i], mask = S.hcap - 1;       // known-set insert: keys of one batch are distinct
            for (uint32_t h = hash_u32(key, mask);; h = (h + 1) & mask) {
                const uint32_t old = atomicCAS(&hk[h], 0xffffffffu, key);
                if (old == 0xffffffffu || old == key) { hv[h] = slot; break; }
            }
            s_flag[i] = 0x100u + rank;
        }
        __syncthreads();
        // warp 0: explore-queue pushes in batch order (container/heap semantics), one warp step per push.  Warps 1-3
        // meanwhile copy the neighbour lists of the fresh vertices, flattened over (vertex, neighbour) so that the loads
        // are independent.
        if (t < 32) {
            const uint32_t nfresh = s_nfresh;
            uint32_t hn = s_heap_n;
            __syncwarp();
            const uint32_t hd_sa = (uint32_t)__cvta_generic_to_shared(s_hd), hs_sa = (uint32_t)__cvta_generic_to_shared(s_hs);
            const uint32_t lt = (1u << t) - 1, le = lt | (1u << t);
            for (uint32_t x0 = 0; x0 < nfresh; x0 += 32) {      // the push list through registers: no load on the push-to-push chain
                const float pd = x0 + t < nfresh ? s_pushd[x0 + t] : 0.f;
                const uint32_t ps = x0 + t < nfresh ? s_pushs[x0 + t] : 0;
                const uint32_t cnt = min(32u, nfresh - x0);
                __syncwarp();
                for (uint32_t x = 0; x < cnt; x++) heap_push_warp(hd_sa, hs_sa, hn++, __shfl_sync(0xffffffffu, pd, x), __shfl_sync(0xffffffffu, ps, x), lt, le);
            }
            if (t == 0) {
                s_heap_n = hn;
                s_known = known0 + nfresh;
                l32[0] = known0 + nfresh;
                l32[2] = this_step + 1;
                l64[2] += B;            // totalQueryNum
                l64[3] += s_succ;       // succQueryNum
                l64[4] += s_server;     // server sub-queries answered for this lane (dummy + successful real)
            }
        } else {
            const uint32_t mv = vec4 ? m / 4 : m, total = s_nfresh * mv;     // 16 bytes per copy where the layout allows
#pragma unroll 4
            for (uint32_t x = t - 32; x < total; x += SR_THREADS - 32) {
                const uint32_t f = x / mv, j = x % mv, r = s_rslot[f], slot = s_pushs[f];
                if (slot >= S.cap) continue;
                const uint32_t *nb = reinterpret_cast<const uint32_t *>(S.cache_entries + ((uint64_t)s_rpart[r] * S.cache_cap + s_rref[r]) * E) + S.dim;
                uint32_t *dst = S.nbr + ((uint64_t)lane * S.cap + slot) * m;
                if (vec4) reinterpret_cast<uint4 *>(dst)[j] = reinterpret_cast<const uint4 *>(nb)[j];
                else dst[j] = nb[j];
            }
        }
        __syncthreads();
    }

    if (next) {
        // search.go:150-167: pop `parallel` vertices (random ids when the queue is empty or in benchmark mode)
        if (t == 0) {
            HeapView hp{s_hd, s_hs, s_heap_n};
            for (uint32_t rept = 0; rept < S.parallel; rept++) s_src[rept] = (hp.n == 0 || benchmarking) ? 0xffffffffu : heap_pop(hp);
            s_heap_n = hp.n;
        }
        __syncthreads();
        const uint64_t rseed = l64[0], rctr0 = l64[1];
        for (uint32_t i = t; i < B; i += SR_THREADS) {
            const uint32_t rept = i / m, j = i % m;
            uint32_t id;
            if (s_src[rept] == 0xffffffffu) {
                uint32_t before = 0;
                for (uint32_t x = 0; x < rept; x++) before += s_src[x] == 0xffffffffu ? m : 0;
                id = (uint32_t)(mix64_dev(rseed, rctr0 + before + j) % (uint64_t)S.n);
            } else {
                id = S.nbr[((uint64_t)lane * S.cap + s_src[rept]) * m + j];
            }
            s_ids[i] = id;
            S.batch_ids[(uint64_t)lane * B + i] = id;
        }
        if (t == 0) {
            uint32_t rnd = 0;
            for (uint32_t x = 0; x < S.parallel; x++) rnd += s_src[x] == 0xffffffffu ? m : 0;
            l64[1] = rctr0 + rnd;
        }
        for (uint32_t p = t; p < S.PN; p += SR_THREADS) s_cnt[p] = 0;
        __syncthreads();
        // batch-pir.go:177-187: bucket by partition in input order; the first `per` of a partition are queried
        uint32_t *s_pi = s_pushs;     // partition of every input (the push list is free again)
        for (uint32_t i = t; i < B; i += SR_THREADS) s_pi[i] = (uint32_t)(s_ids[i] / S.PS);
        __syncthreads();
        for (uint32_t i = t; i < B; i += SR_THREADS) {
            const uint32_t id = s_ids[i], pi = s_pi[i];
            uint32_t rank = 0, first = i;
            for (uint32_t j = 0; j < i; j++) {
                rank += s_pi[j] == pi ? 1u : 0u;
                if (first == i && s_ids[j] == id) first = j;
            }
            s_aux[i] = rank;
            s_flag[i] = first;
            atomicAdd(&s_cnt[pi], 1u);
        }
        __syncthreads();
        for (uint32_t i = t; i < B; i += SR_THREADS) {
            const uint32_t id = s_ids[i], pi = s_pi[i], rank = s_aux[i], first = s_flag[i];
            // every occurrence of an index shares the answer of its first occurrence, if that one was queried at all
            S.batch_rec[(uint64_t)lane * B + i] = (pi < S.PN && s_aux[first] < S.per) ? (int)(pi * S.per + s_aux[first]) : -1;
            if (pi >= S.PN || rank >= S.per) continue;      // surplus in its partition: dropped (zero row)
            const uint32_t gp = lane * S.PN + pi, local = (uint32_t)(id - (uint64_t)pi * S.PS);
            ClientQueryDev q;
            q.part = gp; q.idx = local; q.dummy_seed = 0; q.dummy_ctr = 0;
            if (first != i) {           // pir.go:381-383 via the pending list: the same index earlier in this call
                q.kind = 3;
                q.dummy_seed = pi * S.per + s_aux[first];
            } else {
                const int cs = hash_find(S.cache_key + (uint64_t)gp * S.cache_hcap, S.cache_val + (uint64_t)gp * S.cache_hcap, S.cache_hcap - 1, local);
                if (cs >= 0) { q.kind = 2; q.dummy_seed = (uint64_t)cs; }      // local cache hit: no server query
                else q.kind = 1;
            }
            s_rkind[pi * S.per + rank] = q.kind;
            s_recq[pi * S.per + rank] = q;
        }
        __syncthreads();
        // a real query's answer is finished straight into the next free slot of its sub-PIR's cache (slots of one call:
        // fill count + rank among the call's real records of that part); everything else goes to a scratch row
        for (uint32_t r = t; r < R; r += SR_THREADS) {
            const uint32_t pi = r / S.per, have = min(s_cnt[pi], S.per), gp = lane * S.PN + pi;
            uint64_t *row = S.results + ((uint64_t)a * R + r) * E;
            if (r % S.per < have) {
                ClientQueryDev q = s_recq[r];
                if (q.kind == 1) {
                    uint32_t rank = 0;
                    for (uint32_t x = pi * S.per; x < r; x++) rank += s_rkind[x] == 1 ? 1u : 0u;
                    const uint32_t slot = min(S.cache_cnt[gp] + rank, S.cache_cap - 1);   // cap = MaxQueryNum + 2*per: never clipped
                    q.dummy_seed = slot;
                    row = S.cache_entries + ((uint64_t)gp * S.cache_cap + slot) * E;
                }
                S.records[(uint64_t)a * R + r] = q;
            }
            S.out_rows[(uint64_t)a * R + r] = row;
        }
        // deficit -> dummy queries (batch-pir.go:189-200), SetSize fresh offsets each
        for (uint32_t r = t; r < R; r += SR_THREADS) {
            const uint32_t pi = r / S.per, pos = r % S.per, have = min(s_cnt[pi], S.per);
            if (pos < have) continue;
            const uint32_t gp = lane * S.PN + pi;
            const uint64_t set = S.parts[gp].set_size;
            ClientQueryDev q;
            q.part = gp; q.kind = 0; q.idx = 0;
            q.dummy_seed = S.dummy_seed[gp];
            q.dummy_ctr = S.dummy_ctr[gp] + (uint64_t)(pos - have) * set;
            S.records[(uint64_t)a * R + r] = q;
        }
        __syncthreads();
        for (uint32_t p = t; p < S.PN; p += SR_THREADS) {
            const uint32_t have = min(s_cnt[p], S.per), gp = lane * S.PN + p;
            S.dummy_ctr[gp] += (uint64_t)(S.per - have) * S.parts[gp].set_size;
            uint32_t real = 0;
            for (uint32_t x = p * S.per; x < p * S.per + have; x++) real += s_rkind[x] == 1 ? 1u : 0u;
            if (real) S.cache_cnt[gp] += real;
        }
    }
    __syncthreads();
    heap_n = s_heap_n;
    for (uint32_t i = t * 4; i < heap_n; i += SR_THREADS * 4) {
        *reinterpret_cast<uint4 *>(g_hd + i) = *reinterpret_cast<const uint4 *>(s_hd + i);
        *reinterpret_cast<uint4 *>(g_hs + i) = *reinterpret_cast<const uint4 *>(s_hs + i);
    }
    if (t == 0) l32[1] = heap_n;
}

// search.go:210-233: rank every known vertex by (distance, id), return the k best and the step at which each was reached
__global__ void __launch_bounds__(256) search_final_kernel(SearchDev S, const uint32_t *lane_ids, uint32_t k, long long *ret, long long *step_ret,
                                                           uint64_t *stats, uint64_t *finished) {
    extern __shared__ unsigned long long s_sort[];   // [sort_cap] keys | [sort_cap] slots (u32)
    uint32_t *s_slot = reinterpret_cast<uint32_t *>(s_sort + S.sort_cap);
    const uint32_t lane = lane_ids[blockIdx.x], t = threadIdx.x, cnt = S.lane_u32[(uint64_t)lane * 8];
    for (uint32_t i = t; i < S.sort_cap; i += blockDim.x) {
        if (i < cnt) {
            s_sort[i] = ((unsigned long long)__float_as_uint(S.known_dist[(uint64_t)lane * S.cap + i]) << 32) | S.known_id[(uint64_t)lane * S.cap + i];
            s_slot[i] = i;
        } else {
            s_sort[i] = ~0ull;
            s_slot[i] = 0xffffffffu;
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= S.sort_cap; size <<= 1)
        for (uint32_t stride = size >> 1; stride; stride >>= 1) {
            for (uint32_t i = t; i < S.sort_cap / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = s_sort[lo], y = s_sort[hi];
                if ((x > y) == up) {
                    s_sort[lo] = y; s_sort[hi] = x;
                    const uint32_t sx = s_slot[lo]; s_slot[lo] = s_slot[hi]; s_slot[hi] = sx;
                }
            }
            __syncthreads();
        }
    for (uint32_t i = t; i < k; i += blockDim.x) {
        const bool have = i < cnt;
        ret[(uint64_t)blockIdx.x * k + i] = have ? (long long)(uint32_t)s_sort[i] : -1;
        step_ret[(uint64_t)blockIdx.x * k + i] = have ? (long long)S.known_step[(uint64_t)lane * S.cap + s_slot[i]] : -1;
    }
    if (t < 3) stats[(uint64_t)blockIdx.x * 3 + t] = S.lane_u64[(uint64_t)lane * 8 + 2 + t];
    for (uint32_t p = t; p < S.PN; p += blockDim.x) finished[(uint64_t)blockIdx.x * S.PN + p] = *S.parts[lane * S.PN + p].finished;
}

__global__ void search_fill_u32_kernel(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace pm

// =============================================================================================
// C-ABI: pm_search_*
// =============================================================================================
struct pm_search {
    pm_client *c;
    pm_search_config cfg;
    pm::SearchDev D;
    void *arena = nullptr, *call = nullptr;
    uint64_t L = 0, act = 0, k = 0;
    int benchmarking = 0;
    uint32_t *d_lanes = nullptr, *d_part_map = nullptr, *d_vid = nullptr;   // [L], [L*PN], [L*R]
    uint64_t *d_rseed = nullptr;
    // per-call device buffers (sized for L lanes)
    uint32_t *d_off = nullptr, *d_chunk = nullptr, *d_set = nullptr;
    uint64_t *d_row0 = nullptr, *d_nrows = nullptr, *d_ans = nullptr, *d_res = nullptr;
    pm::ClientMeta *d_meta = nullptr;
    float *d_dist = nullptr;
    long long *d_ret = nullptr, *d_step = nullptr;
    uint64_t *d_stats = nullptr, *d_fin = nullptr;
    void *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    uint64_t stride = 0;
    // the HBM-bound answer kernel of a step runs on a LOW-priority stream of its own (ordered by events): the prepare / step /
    // finish kernels of another group, on their high-priority client stream, then interleave with it instead of queueing
    cudaStream_t ans_stream = nullptr;
    cudaEvent_t ev_prep = nullptr, ev_ans = nullptr;
    size_t step_smem = 0, begin_smem = 0, final_smem = 0;
};

namespace pm {
static uint32_t next_pow2(uint64_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
// the parts of a client whose hint tables are rebuilt start with an empty local cache (pir.go:203-255 / Initialization)
int search_clear_cache(pm_search *s, const uint32_t *part_ids, uint64_t n, cudaStream_t st) {
    for (uint64_t i = 0; i < n; i++) {
        const uint32_t gp = part_ids[i];
        search_fill_u32_kernel<<<16, 256, 0, st>>>(s->D.cache_key + (uint64_t)gp * s->D.cache_hcap, s->D.cache_hcap, 0xffffffffu);
        PM_CUDA(cudaMemsetAsync(s->D.cache_cnt + gp, 0, 4, st));
        count_launch();
    }
    PM_CHECK_LAUNCH();
    return PM_OK;
}
}  // namespace pm

PM_EXPORT int pm_search_create(pm_client *c, const pm_search_config *cfg, pm_search **out) {
    using namespace pm;
    if (!c || !cfg || !out) return set_error(PM_ERR_ARG, "pm_search_create: null pointer");
    *out = nullptr;
    const uint64_t L = cfg->lanes, PN = cfg->partition_num, B = cfg->parallel * cfg->m;
    if (L == 0 || PN == 0 || L * PN != c->n_parts) return set_error(PM_ERR_ARG, "pm_search_create: lanes * partition_num must equal the client's parts");
    if ((cfg->dim + cfg->m) % 2 || (cfg->dim + cfg->m) / 2 != c->E) return set_error(PM_ERR_ARG, "pm_search_create: entry is not dim f32 || m u32");
    if (B == 0 || B / PN == 0 || cfg->n == 0 || cfg->n > 0x7fffffffull || cfg->n > c->db->n_rows || cfg->max_step == 0 || cfg->n_start == 0)
        return set_error(PM_ERR_ARG, "pm_search_create: bad search geometry");
    if (B > 4096 || cfg->parallel > 64) return set_error(PM_ERR_UNSUPPORTED, "pm_search_create: parallel * m too large");
    int rc = ensure_device(c->db->device);
    if (rc) return rc;
    pm_search *s = new (std::nothrow) pm_search();
    if (!s) return set_error(PM_ERR_NOMEM, "out of host memory");
    s->c = c; s->cfg = *cfg; s->L = L;
    SearchDev &D = s->D;
    memset(&D, 0, sizeof(D));
    D.n = (uint32_t)cfg->n; D.dim = (uint32_t)cfg->dim; D.m = (uint32_t)cfg->m; D.E = (uint32_t)c->E; D.PN = (uint32_t)PN;
    D.per = (uint32_t)(B / PN); D.B = (uint32_t)B; D.R = D.PN * D.per; D.ns = (uint32_t)cfg->n_start;
    D.cap = (uint32_t)((B * cfg->max_step + cfg->parallel + 3) & ~3ull);
    D.hcap = next_pow2(2ull * D.cap); D.sort_cap = next_pow2(D.cap);
    // a slot is handed out per real query, successful or not: at most `per` a call, and a batch epoch ends before
    // MaxQueryNum + per of them (batch-pir.go:239-245)
    D.cache_cap = (uint32_t)(std::max<uint64_t>(cfg->cache_entries, 1) + 2 * (B / PN) + 8); D.cache_hcap = next_pow2(2ull * D.cache_cap);
    D.max_step = (uint32_t)cfg->max_step; D.parallel = (uint32_t)cfg->parallel; D.PS = cfg->partition_size;
    D.db = c->db->d_rows; D.parts = c->d_parts;
    if (D.sort_cap > 8192) { delete s; return set_error(PM_ERR_UNSUPPORTED, "pm_search_create: max_step * parallel * m too large for the final sort"); }
    // one arena for everything that lives as long as the object
    size_t bytes = 0;
    auto take = [&](size_t b) { size_t o = bytes; bytes += (b + 255) & ~(size_t)255; return o; };
    const uint64_t NPARTS = L * PN;
    const size_t o_l64 = take(L * 64), o_l32 = take(L * 32), o_kid = take(L * D.cap * 4), o_kstep = take(L * D.cap * 4), o_kdist = take(L * D.cap * 4),
                 o_hk = take(L * D.hcap * 4), o_hv = take(L * D.hcap * 4), o_hd = take(L * D.cap * 4), o_hs = take(L * D.cap * 4),
                 o_nbr = take(L * D.cap * D.m * 4), o_bid = take(L * B * 4), o_brec = take(L * B * 4), o_q = take(L * D.dim * 4),
                 o_sid = take(L * D.ns * 4), o_snb = take(L * D.ns * D.m * 4), o_sv = take(L * D.ns * D.dim * 4), o_sd = take(L * D.ns * 4),
                 o_ck = take(NPARTS * D.cache_hcap * 4), o_cv = take(NPARTS * D.cache_hcap * 4), o_cc = take(NPARTS * 4),
                 o_ce = take(NPARTS * (size_t)D.cache_cap * D.E * 8), o_ds = take(NPARTS * 8), o_dc = take(NPARTS * 8),
                 o_lanes = take(L * 4), o_pmap = take(NPARTS * 4), o_vid = take(L * D.R * 4), o_rs = take(L * 8);
    // per-call buffers
    s->stride = (c->max_set + 3) & ~3ull;
    const uint64_t Q = L * D.R;
    const size_t o_rec = take(Q * sizeof(ClientQueryDev)), o_off = take(Q * s->stride * 4), o_row0 = take(Q * 8), o_nrows = take(Q * 8),
                 o_chunk = take(Q * 4), o_set = take(Q * 4), o_ans = take(Q * D.E * 8), o_res = take(Q * D.E * 8), o_meta = take(Q * sizeof(ClientMeta)),
                 o_dist = take(Q * 4), o_ret = take(L * 8 * 4096), o_stats = take(L * 24), o_fin = take(NPARTS * 8), o_rows = take(Q * 8);
    cudaError_t e = cudaMalloc(&s->arena, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); delete s; return set_error(PM_ERR_NOMEM, "pm_search_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    char *base = (char *)s->arena;
    D.lane_u64 = (uint64_t *)(base + o_l64); D.lane_u32 = (uint32_t *)(base + o_l32);
    D.known_id = (uint32_t *)(base + o_kid); D.known_step = (uint32_t *)(base + o_kstep); D.known_dist = (float *)(base + o_kdist);
    D.hash_key = (uint32_t *)(base + o_hk); D.hash_val = (uint32_t *)(base + o_hv); D.heap_dist = (float *)(base + o_hd); D.heap_slot = (uint32_t *)(base + o_hs);
    D.nbr = (uint32_t *)(base + o_nbr); D.batch_ids = (uint32_t *)(base + o_bid); D.batch_rec = (int32_t *)(base + o_brec); D.qvec = (float *)(base + o_q);
    D.start_id = (uint32_t *)(base + o_sid); D.start_nbr = (uint32_t *)(base + o_snb); D.start_vec = (float *)(base + o_sv); D.start_dist = (float *)(base + o_sd);
    D.cache_key = (uint32_t *)(base + o_ck); D.cache_val = (uint32_t *)(base + o_cv); D.cache_cnt = (uint32_t *)(base + o_cc);
    D.cache_entries = (uint64_t *)(base + o_ce); D.dummy_seed = (uint64_t *)(base + o_ds); D.dummy_ctr = (uint64_t *)(base + o_dc);
    s->d_lanes = (uint32_t *)(base + o_lanes); s->d_part_map = (uint32_t *)(base + o_pmap); s->d_vid = (uint32_t *)(base + o_vid); s->d_rseed = (uint64_t *)(base + o_rs);
    D.records = (ClientQueryDev *)(base + o_rec);
    s->d_off = (uint32_t *)(base + o_off); s->d_row0 = (uint64_t *)(base + o_row0); s->d_nrows = (uint64_t *)(base + o_nrows);
    s->d_chunk = (uint32_t *)(base + o_chunk); s->d_set = (uint32_t *)(base + o_set); s->d_ans = (uint64_t *)(base + o_ans); s->d_res = (uint64_t *)(base + o_res);
    s->d_meta = (ClientMeta *)(base + o_meta); s->d_dist = (float *)(base + o_dist);
    s->d_ret = (long long *)(base + o_ret); s->d_step = s->d_ret + L * 2048; s->d_stats = (uint64_t *)(base + o_stats); s->d_fin = (uint64_t *)(base + o_fin);
    D.meta = s->d_meta; D.results = s->d_res; D.dist = s->d_dist; D.out_rows = (uint64_t **)(base + o_rows);
    std::lock_guard<std::mutex> lock(c->mu);
    e = cudaMemsetAsync(s->arena, 0, o_rec, c->stream);     // counters, dummy counters, caches counts, state
    if (e == cudaSuccess) {
        search_fill_u32_kernel<<<256, 256, 0, c->stream>>>(D.cache_key, NPARTS * D.cache_hcap, 0xffffffffu);
        count_launch();
        std::vector<uint32_t> vid(L * D.R);
        for (uint64_t i = 0; i < vid.size(); i++) vid[i] = (uint32_t)(i / D.R);
        e = cudaMemcpyAsync(s->d_vid, vid.data(), vid.size() * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    if (e != cudaSuccess) { cudaFree(s->arena); delete s; return set_error(PM_ERR_CUDA, "pm_search_create: %s", cudaGetErrorString(e)); }
    D.recq_off = (uint32_t)(((size_t)D.cap * 8 + (size_t)B * 24 + (size_t)D.R * 24 + D.parallel * 4 + PN * 4 + 63) & ~(size_t)63);
    s->step_smem = D.recq_off + (size_t)D.R * sizeof(ClientQueryDev) + 64;
    s->begin_smem = (size_t)D.ns * 8;
    s->final_smem = (size_t)D.sort_cap * 12;
    if (s->step_smem > 200 * 1024 || s->begin_smem > 200 * 1024) { cudaFree(s->arena); delete s; return set_error(PM_ERR_UNSUPPORTED, "pm_search_create: search state too large for shared memory"); }
    cudaFuncSetAttribute(search_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->step_smem);
    cudaFuncSetAttribute(search_begin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->begin_smem);
    cudaFuncSetAttribute(search_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->final_smem);
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (tune(T_SEARCH_ANS_STREAM) != 0 &&
            (cudaStreamCreateWithPriority(&s->ans_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
             cudaEventCreateWithFlags(&s->ev_prep, cudaEventDisableTiming) != cudaSuccess ||
             cudaEventCreateWithFlags(&s->ev_ans, cudaEventDisableTiming) != cudaSuccess)) {
            cudaGetLastError();
            s->ans_stream = nullptr;     // fall back to one stream
        }
    }
    c->search = s;
    *out = s;
    return PM_OK;
}

PM_EXPORT int pm_search_destroy(pm_search *s) {
    if (!s) return PM_OK;
    if (pm::ensure_device(s->c->db->device) == PM_OK) {
        cudaStreamSynchronize(s->c->stream);
        if (s->ans_stream) { cudaStreamSynchronize(s->ans_stream); cudaStreamDestroy(s->ans_stream); }
        if (s->ev_prep) cudaEventDestroy(s->ev_prep);
        if (s->ev_ans) cudaEventDestroy(s->ev_ans);
        cudaFree(s->arena);
        if (s->h_stage) cudaFreeHost(s->h_stage);
    }
    if (s->c->search == s) s->c->search = nullptr;
    delete s;
    return PM_OK;
}

PM_EXPORT int pm_search_set_start(pm_search *s, uint32_t lane, const int64_t *ids, const float *vectors, const int32_t *neighbors) {
    using namespace pm;
    if (!s || !ids || !vectors || !neighbors || lane >= s->L) return set_error(PM_ERR_ARG, "pm_search_set_start: bad argument");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    const SearchDev &D = s->D;
    std::vector<uint32_t> id32(D.ns);
    for (uint32_t i = 0; i < D.ns; i++) id32[i] = (uint32_t)ids[i];
    std::lock_guard<std::mutex> lock(s->c->mu);
    PM_CUDA(cudaMemcpyAsync(D.start_id + (uint64_t)lane * D.ns, id32.data(), D.ns * 4, cudaMemcpyHostToDevice, s->c->stream));
    PM_CUDA(cudaMemcpyAsync(D.start_vec + (uint64_t)lane * D.ns * D.dim, vectors, (size_t)D.ns * D.dim * 4, cudaMemcpyHostToDevice, s->c->stream));
    PM_CUDA(cudaMemcpyAsync(D.start_nbr + (uint64_t)lane * D.ns * D.m, neighbors, (size_t)D.ns * D.m * 4, cudaMemcpyHostToDevice, s->c->stream));
    PM_CUDA(cudaStreamSynchronize(s->c->stream));
    return PM_OK;
}

PM_EXPORT int pm_search_set_dummy_seed(pm_search *s, uint32_t lane, const uint64_t *seeds) {
    using namespace pm;
    if (!s || !seeds || lane >= s->L) return set_error(PM_ERR_ARG, "pm_search_set_dummy_seed: bad argument");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(s->c->mu);
    PM_CUDA(cudaMemcpyAsync(s->D.dummy_seed + (uint64_t)lane * s->D.PN, seeds, s->D.PN * 8, cudaMemcpyHostToDevice, s->c->stream));
    PM_CUDA(cudaStreamSynchronize(s->c->stream));
    return PM_OK;
}

PM_EXPORT int pm_search_begin(pm_search *s, const uint32_t *lanes, uint64_t act, const float *queries, const uint64_t *rand_seeds, uint64_t k,
                              int benchmarking) {
    using namespace pm;
    if (!s || !lanes || !queries || !rand_seeds || act == 0 || act > s->L || k == 0 || k > 2048) return set_error(PM_ERR_ARG, "pm_search_begin: bad argument");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    const SearchDev &D = s->D;
    pm_client *c = s->c;
    std::lock_guard<std::mutex> lock(c->mu);
    // staging (page-locked): lane ids | part map | random seeds | query vectors
    const size_t b_l = act * 4, b_pm = act * D.PN * 4, b_rs = act * 8, b_q = act * D.dim * 4;
    const size_t need = b_l + b_pm + b_rs + b_q + std::max<size_t>(act * k * 16 + act * 24 + act * D.PN * 8, 4096);
    if (s->h_stage_bytes < need) {
        if (s->h_stage) { PM_CUDA(cudaStreamSynchronize(c->stream)); cudaFreeHost(s->h_stage); }
        s->h_stage = nullptr; s->h_stage_bytes = 0;
        PM_CUDA(cudaHostAlloc(&s->h_stage, need * 2, cudaHostAllocDefault));
        s->h_stage_bytes = need * 2;
    }
    PM_CUDA(cudaStreamSynchronize(c->stream));   // the staging block of the previous round is free again
    char *h = (char *)s->h_stage;
    uint32_t *hl = (uint32_t *)h, *hpm = (uint32_t *)(h + b_l);
    uint64_t *hrs = (uint64_t *)(h + b_l + b_pm);
    for (uint64_t a = 0; a < act; a++) {
        if (lanes[a] >= s->L) return set_error(PM_ERR_ARG, "pm_search_begin: lane out of range");
        hl[a] = lanes[a];
        for (uint32_t p = 0; p < D.PN; p++) hpm[a * D.PN + p] = lanes[a] * D.PN + p;
        hrs[a] = rand_seeds[a];
    }
    PM_CUDA(cudaMemcpyAsync(s->d_lanes, hl, b_l, cudaMemcpyHostToDevice, c->stream));
    PM_CUDA(cudaMemcpyAsync(s->d_part_map, hpm, b_pm, cudaMemcpyHostToDevice, c->stream));
    PM_CUDA(cudaMemcpyAsync(s->d_rseed, hrs, b_rs, cudaMemcpyHostToDevice, c->stream));
    float *hq = (float *)(h + b_l + b_pm + b_rs);
    memcpy(hq, queries, b_q);
    PM_CUDA(cudaMemcpyAsync(D.qvec, hq, b_q, cudaMemcpyHostToDevice, c->stream));   // indexed by position in the round
    s->act = act; s->k = k; s->benchmarking = benchmarking;
    if (!benchmarking) {
        dim3 grid((unsigned)act, (unsigned)std::min<uint32_t>(32, (D.ns + 63) / 64));
        search_start_dist_kernel<<<grid, SR_THREADS, 0, c->stream>>>(D, s->d_lanes);
        count_launch();
    }
    search_begin_kernel<<<(unsigned)act, SR_THREADS, s->begin_smem, c->stream>>>(D, s->d_lanes, s->d_rseed, benchmarking ? 1u : 0u);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}

namespace pm {
static int search_step_launch(pm_search *s, bool apply, bool next) {
    search_step_kernel<<<(unsigned)s->act, SR_THREADS, s->step_smem, s->c->stream>>>(s->D, s->d_lanes, apply ? 1u : 0u, next ? 1u : 0u, s->benchmarking ? 1u : 0u);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
}  // namespace pm

// One search step for every lane of the round: [apply the previous fetch,] pop + next batch + records, then the PIR fetch
// (prepare -> answer -> finish with distances).  Enqueue only.
PM_EXPORT int pm_search_fetch(pm_search *s, int apply_previous) {
    using namespace pm;
    if (!s || s->act == 0) return set_error(PM_ERR_ARG, "pm_search_fetch: no round in progress");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    pm_client *c = s->c;
    const SearchDev &D = s->D;
    std::lock_guard<std::mutex> lock(c->mu);
    if ((rc = search_step_launch(s, apply_previous != 0, true))) return rc;
    const uint64_t q = s->act * D.R, nparts = s->act * D.PN;
    uint64_t max_p = 0;
    for (uint64_t i = 0; i < c->n_parts; i++) if (c->host_parts[i].poff) max_p = std::max<uint64_t>(max_p, c->host_parts[i].n_primary);
    const size_t smem_base = (aes_tab_words<8>() + 64 + s->stride + CL_MAX_LIST) * 4;
    const uint32_t mirror = smem_base + max_p * 4 <= 200 * 1024 ? 1u : 0u;
    const size_t smem = smem_base + (mirror ? max_p * 4 : 0);
    if (smem > 200 * 1024) return set_error(PM_ERR_UNSUPPORTED, "pm_search_fetch: set_size too large for the prepare kernel's shared memory");
    PM_CUDA(cudaFuncSetAttribute(client_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned prep_threads = nparts <= 64 ? CL_THREADS : CL_THREADS / 2;
    client_prepare_kernel<<<(unsigned)nparts, prep_threads, smem, c->stream>>>(c->d_parts, D.records, nullptr, nullptr, (uint32_t)q, (uint32_t)s->stride, s->d_off,
                                                                                s->d_meta, s->d_row0, s->d_nrows, s->d_chunk, s->d_set, mirror, s->d_part_map, D.per);
    PM_CHECK_LAUNCH();
    count_launch();
    if (s->ans_stream) {
        PM_CUDA(cudaEventRecord(s->ev_prep, c->stream));
        PM_CUDA(cudaStreamWaitEvent(s->ans_stream, s->ev_prep, 0));
        if ((rc = answer_enqueue(c->db, s->d_row0, s->d_nrows, s->d_chunk, s->d_set, s->d_off, s->stride, q, (uint32_t)s->stride, s->d_ans, s->ans_stream))) return rc;
        PM_CUDA(cudaEventRecord(s->ev_ans, s->ans_stream));
        PM_CUDA(cudaStreamWaitEvent(c->stream, s->ev_ans, 0));
    } else if ((rc = answer_enqueue(c->db, s->d_row0, s->d_nrows, s->d_chunk, s->d_set, s->d_off, s->stride, q, (uint32_t)s->stride, s->d_ans, c->stream))) {
        return rc;
    }
    client_finish_kernel<<<(unsigned)nparts, 256, 0, c->stream>>>(c->d_parts, nullptr, nullptr, s->d_meta, D.E, s->d_ans, s->d_res,
                                                                  s->benchmarking ? nullptr : D.qvec, s->d_vid, D.dim, s->benchmarking ? nullptr : s->d_dist,
                                                                  s->d_part_map, D.per, D.out_rows);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}

// apply the last fetch without starting another one (before a re-preprocessing, which clears the caches the next batch
// must see cleared: batch-pir.go:239-245 runs after the responses of the call have been booked)
PM_EXPORT int pm_search_apply(pm_search *s) {
    using namespace pm;
    if (!s || s->act == 0) return set_error(PM_ERR_ARG, "pm_search_apply: no round in progress");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(s->c->mu);
    return search_step_launch(s, true, false);
}

PM_EXPORT int pm_search_finish(pm_search *s, int apply_previous, int64_t *ret, int64_t *step_ret, uint64_t *stats, uint64_t *finished) {
    using namespace pm;
    if (!s || s->act == 0 || !ret || !step_ret || !stats || !finished) return set_error(PM_ERR_ARG, "pm_search_finish: bad argument");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    pm_client *c = s->c;
    const SearchDev &D = s->D;
    std::lock_guard<std::mutex> lock(c->mu);
    if (apply_previous && (rc = search_step_launch(s, true, false))) return rc;
    const uint64_t act = s->act, k = s->k;
    search_final_kernel<<<(unsigned)act, 256, s->final_smem, c->stream>>>(D, s->d_lanes, (uint32_t)k, s->d_ret, s->d_step, s->d_stats, s->d_fin);
    PM_CHECK_LAUNCH();
    count_launch();
    char *h = (char *)s->h_stage;
    const size_t b_ret = act * k * 8, b_st = act * 24, b_fin = act * D.PN * 8;
    PM_CUDA(cudaMemcpyAsync(h, s->d_ret, b_ret, cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaMemcpyAsync(h + b_ret, s->d_step, b_ret, cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaMemcpyAsync(h + 2 * b_ret, s->d_stats, b_st, cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaMemcpyAsync(h + 2 * b_ret + b_st, s->d_fin, b_fin, cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(ret, h, b_ret);
    memcpy(step_ret, h + b_ret, b_ret);
    memcpy(stats, h + 2 * b_ret, b_st);
    memcpy(finished, h + 2 * b_ret + b_st, b_fin);
    s->act = 0;
    return PM_OK;
}

// the local cache of one sub-PIR of one lane, for a lane that leaves the device path (tests, checkpointing)
PM_EXPORT int pm_search_cache_download(pm_search *s, uint32_t lane, uint32_t part, uint64_t *idx_out, uint64_t *entries_out, uint64_t cap,
                                       uint64_t *count) {
    using namespace pm;
    if (!s || !count || lane >= s->L || part >= s->D.PN) return set_error(PM_ERR_ARG, "pm_search_cache_download: bad argument");
    int rc = ensure_device(s->c->db->device);
    if (rc) return rc;
    const SearchDev &D = s->D;
    const uint64_t gp = (uint64_t)lane * D.PN + part;
    std::lock_guard<std::mutex> lock(s->c->mu);
    uint32_t cnt = 0;
    PM_CUDA(cudaMemcpyAsync(&cnt, D.cache_cnt + gp, 4, cudaMemcpyDeviceToHost, s->c->stream));
    PM_CUDA(cudaStreamSynchronize(s->c->stream));
    // slots are handed out per query and only successful ones get a key: count and compact by walking the key table
    std::vector<uint32_t> keys(D.cache_hcap), vals(D.cache_hcap);
    PM_CUDA(cudaMemcpyAsync(keys.data(), D.cache_key + gp * D.cache_hcap, D.cache_hcap * 4, cudaMemcpyDeviceToHost, s->c->stream));
    PM_CUDA(cudaMemcpyAsync(vals.data(), D.cache_val + gp * D.cache_hcap, D.cache_hcap * 4, cudaMemcpyDeviceToHost, s->c->stream));
    PM_CUDA(cudaStreamSynchronize(s->c->stream));
    uint64_t n_valid = 0;
    for (uint32_t h = 0; h < D.cache_hcap; h++) n_valid += keys[h] != 0xffffffffu && vals[h] < cnt;
    *count = n_valid;
    if (n_valid == 0 || !idx_out || !entries_out) return PM_OK;
    if (cap < n_valid) return set_error(PM_ERR_ARG, "pm_search_cache_download: buffer too small");
    std::vector<uint64_t> slab((size_t)cnt * D.E);
    PM_CUDA(cudaMemcpyAsync(slab.data(), D.cache_entries + gp * D.cache_cap * D.E, (size_t)cnt * D.E * 8, cudaMemcpyDeviceToHost, s->c->stream));
    PM_CUDA(cudaStreamSynchronize(s->c->stream));
    uint64_t o = 0;
    for (uint32_t h = 0; h < D.cache_hcap; h++) {
        if (keys[h] == 0xffffffffu || vals[h] >= cnt) continue;
        idx_out[o] = keys[h];
        memcpy(entries_out + o * D.E, slab.data() + (size_t)vals[h] * D.E, D.E * 8);
        o++;
    }
    return PM_OK;
}
