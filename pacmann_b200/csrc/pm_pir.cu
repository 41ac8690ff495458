// PianoPIR kernels for sm_100a: AES-PRF, key schedule, hint generation (A5/A7), server answer
// (A6/A8), row gather, xorSlices.  See include/pacmann_cuda.h for the reference seams.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pm_aes.cuh"
#include <cooperative_groups.h>

#include "pm_common.cuh"

namespace cg = cooperative_groups;

namespace pm {

__constant__ uint32_t c_te0[256];

// ---------------------------------------------------------------------------------------------
// host-side table construction (S-box computed from GF(2^8) inversion + affine map, FIPS-197 5.1.1)
// ---------------------------------------------------------------------------------------------
static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) p ^= a;
        uint8_t hi = a & 0x80;
        a = (uint8_t)(a << 1);
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}
static void build_te0(uint32_t te0[256]) {
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        for (int y = 1; y < 256 && x; y++)
            if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        s ^= 0x63;
        te0[x] = (uint32_t)gf_mul(s, 2) | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)gf_mul(s, 3) << 24);
    }
}
int upload_tables() {  // called once per device from ensure_device()
    uint32_t te0[256];
    build_te0(te0);
    PM_CUDA(cudaMemcpyToSymbol(c_te0, te0, sizeof(te0)));
    return PM_OK;
}

// ---------------------------------------------------------------------------------------------
// A1 key schedule, A2 generic PRF
// ---------------------------------------------------------------------------------------------
struct RkArray { uint32_t w[44]; };
struct RkOfArray {
    const RkArray &a;
    __device__ __forceinline__ uint32_t operator[](int i) const { return a.w[i]; }
};

__global__ void expand_key_kernel(const uint32_t *key_words, uint32_t *rk_out, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t *key = key_words + 4 * t;
    uint32_t *rk = rk_out + 44 * t;
    uint32_t w[44];
    for (int i = 0; i < 4; i++) w[i] = key[i];
    uint32_t rcon = 1;
    for (int i = 4; i < 44; i++) {
        uint32_t t = w[i - 1];
        if ((i & 3) == 0) {
            t = (t >> 8) | (t << 24);  // RotWord on LE-packed bytes
            t = ((c_te0[t & 0xff] >> 8) & 0xff) | (((c_te0[(t >> 8) & 0xff] >> 8) & 0xff) << 8) |
                (((c_te0[(t >> 16) & 0xff] >> 8) & 0xff) << 16) | (((c_te0[t >> 24] >> 8) & 0xff) << 24);
            t ^= rcon;
            rcon = (rcon << 1) ^ ((rcon & 0x80) ? 0x11b : 0);
        }
        w[i] = w[i - 4] ^ t;
    }
    for (int i = 0; i < 44; i++) rk[i] = w[i];
}

__global__ void __launch_bounds__(256) prf_batch_kernel(const __grid_constant__ RkArray rk, const uint64_t *tags,
                                                        const uint64_t *xs, uint64_t n, uint64_t *out) {
    extern __shared__ uint32_t smem[];
    aes_tab_fill<1>(smem, c_te0);
    __syncthreads();
    AesTab<1> T{smem + (threadIdx.x & 31)};
    RkOfArray R{rk};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = (tags[i] << 35) + xs[i];
        uint32_t in0 = (uint32_t)v, in1 = (uint32_t)(v >> 32);
        uint32_t s0 = in0, s1 = in1, s2 = 0, s3 = 0;
        aes128_encrypt<1>(T, R, s0, s1, s2, s3);
        out[i] = (uint64_t)(s0 ^ in0) | ((uint64_t)(s1 ^ in1) << 32);
    }
}

// ---------------------------------------------------------------------------------------------
// A5/A7 hint generation
// ---------------------------------------------------------------------------------------------
// Hint-stationary: a group of G lanes owns one hint and keeps its parity in registers across all
// set_size chunks; each lane of the group evaluates the PRF for a different chunk (G chunks per
// step, all 32 lanes of the warp busy in AES), offsets are exchanged by shuffle, and the G lanes
// then read the selected row together (G*16 contiguous bytes per load instruction, whole 128-byte
// lines for G = 8).  Parities are written once.  CTAs walk tiles in a static round-robin so that
// all resident CTAs sweep the same DB slice chunk 0..S-1 at the same pace: each chunk is pulled
// from HBM once and re-hit in L2 by the other ~H'/ChunkSize hints that select rows of it.
constexpr int HG_THREADS = 512;        // default CTA width (16 warps); the width is chosen per launch: 12..16 warps
constexpr int HG_MAX_THREADS = 512;    // wider CTAs would cap the kernel below the 128 registers it needs (spills)
constexpr int HG_MAX_JOBS = 16;

struct HintJobDev {
    uint32_t rk[44];
    uint64_t row0, n_rows;
    uint64_t hint_begin, n_hints, n_primary, backup_group;
    const uint64_t *tags;
    const int32_t *skip;
    uint64_t *out;
    uint32_t chunk_mask, chunk_shift, set_size, tile_begin;
};
struct HintParams {
    HintJobDev jobs[HG_MAX_JOBS];
    const void *db;
    uint32_t n_jobs, n_tiles;
    uint32_t threads;   // CTA width of this launch
    uint32_t ev, evx;  // vectors per row (incl. un-xored tail), vectors that are xored
    unsigned int *sync;  // round barrier counter (cooperative launch) or nullptr
};
struct RkOfJob {
    const HintParams &P;
    int j;
    __device__ __forceinline__ uint32_t operator[](int i) const { return P.jobs[j].rk[i]; }
};

__device__ __forceinline__ uint4 ldg_row(const uint4 *p, bool pred) {
    uint4 v;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\tmov.u32 %2, 0;\n\tmov.u32 %3, 0;\n\t"
        "@p ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w)
        : "l"(p), "r"((uint32_t)pred));
    return v;
}
__device__ __forceinline__ uint2 ldg_row(const uint2 *p, bool pred) {
    uint2 v;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.u32 %0, 0;\n\tmov.u32 %1, 0;\n\t"
        "@p ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];\n\t}"
        : "=&r"(v.x), "=&r"(v.y)
        : "l"(p), "r"((uint32_t)pred));
    return v;
}

// ---- software pipelining and load policy ------------------------------------------------------------
// Details of the inner loop, all about keeping the pipes busy:
//  * the PRF of the NEXT group of G chunks is evaluated in NPH = G/U slices interleaved with the U-row load
//    batches of the CURRENT group, so each warp hides its own row-load latency behind its own AES work;
//  * no per-load predicates: a (hint, chunk) pair that must not contribute (zero padding past n_rows, the
//    skipped chunk, inactive lanes, chunks past S in the last group) still loads a real, harmless row -- a
//    different one per pair, so no L2 line becomes a hot spot -- and bit 31 of the exchanged row word says
//    whether to XOR it.  The common case (every pair of the warp's batch contributes) is one warp-uniform
//    vote and the unmasked 3-input XOR;
//  * FULL = the row is an exact multiple of G vectors (896 B, 640 B, 128 B rows): no column predicates either.
constexpr uint32_t ROW_CONTRIB = 0x80000000u;
template <typename VT, int G, int NV, int NTAB, int NB, int U, bool FULL, int PH, int NPH, typename RK>
__device__ __forceinline__ void hg_phase(const AesTab<NTAB> &T, const RK &R, const PrfTagPart &g, uint32_t c_next,
                                         PrfState &st, uint32_t row_cur, int gbase, int gl, const VT *base,
                                         uint32_t ev, uint32_t evx, VT (&par)[NV]) {
    VT buf[U][NV];
    uint32_t rr[U];
    bool all_in = true;
#pragma unroll
    for (int u = 0; u < U; u++) {
        rr[u] = __shfl_sync(0xffffffffu, row_cur, gbase + PH * U + u);
        all_in = all_in && (rr[u] & ROW_CONTRIB);
        const VT *rp = base + (uint64_t)(rr[u] & ~ROW_CONTRIB) * ev;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            if (FULL) buf[u][k] = ldg_stream(rp + k * G);
            else buf[u][k] = ldg_row(rp + k * G, (uint32_t)(k * G + gl) < evx);
        }
    }
    prf_rounds<NTAB, NB, PrfPhase<PH, NPH>::first, PrfPhase<PH, NPH>::last>(T, R, g, c_next, st);
    if (__all_sync(0xffffffffu, all_in)) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int k = 0; k < NV; k++) vxor(par[k], buf[u][k]);
    } else {
#pragma unroll
        for (int u = 0; u < U; u++)
            if (rr[u] & ROW_CONTRIB) {
#pragma unroll
                for (int k = 0; k < NV; k++) vxor(par[k], buf[u][k]);
            }
    }
}

template <typename VT, int G, int NV, int NTAB, int NB, int U, bool FULL>
__global__ void __launch_bounds__(HG_MAX_THREADS, 1) hintgen_kernel(const __grid_constant__ HintParams P) {
    extern __shared__ uint32_t smem[];
    aes_tab_fill<NTAB>(smem, c_te0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const AesTab<NTAB> T{smem + lane};
    constexpr int GPW = 32 / G, NPH = G / U;
    const int gl = lane & (G - 1), gbase = lane & ~(G - 1), gw = lane / G;
    const uint32_t hints_per_tile = (blockDim.x / 32) * GPW;
    const uint32_t ev = P.ev, evx = P.evx;

    // Rounds: in round r the CTAs process tiles r*grid .. r*grid+grid-1, i.e. one contiguous range of hints of
    // (nearly always) one sub-PIR, and all of them sweep that sub-PIR's DB slice from chunk 0.  With the round
    // barrier the CTAs start every sweep together, so a chunk is fetched from HBM once per round and re-hit in L2
    // by the other CTAs; without it they drift apart until the sweeps are uncorrelated.
    const uint32_t rounds = (P.n_tiles + gridDim.x - 1) / gridDim.x;
    for (uint32_t round = 0; round < rounds; round++) {
        const uint32_t tile = round * gridDim.x + blockIdx.x;
        if (P.sync && round > 0) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(P.sync, 1u);
                const unsigned int target = round * gridDim.x;
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.sync));
                    if (seen < target) __nanosleep(100);
                } while (seen < target);
            }
            __syncthreads();
        }
        if (tile >= P.n_tiles) continue;
        int j = 0;
        while (j + 1 < (int)P.n_jobs && tile >= P.jobs[j + 1].tile_begin) j++;
        const HintJobDev &J = P.jobs[j];
        const RkOfJob R{P, j};
        const uint64_t i = (uint64_t)(tile - J.tile_begin) * hints_per_tile + warp * GPW + gw;
        const bool active = i < J.n_hints;
        uint64_t tag = 0;
        int32_t skip = PM_NO_SKIP;
        if (active) {
            const uint64_t h = J.hint_begin + i;
            tag = J.tags ? J.tags[i] : h;
            if (J.skip) skip = J.skip[i];
            else if (h >= J.n_primary && J.backup_group) skip = (int32_t)((h - J.n_primary) / J.backup_group);
        }
        const PrfTagPart g = prf_tag_part(T, R, tag);
        const uint32_t S = J.set_size, cmask = J.chunk_mask, cshift = J.chunk_shift;
        const uint32_t n_rows = (uint32_t)J.n_rows;
        const VT *base = reinterpret_cast<const VT *>(P.db) + J.row0 * ev + gl;

        VT par[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) vzero(par[k]);

        // row word exchanged inside the group: bits 0..30 = a row that is always safe to read, bit 31 = contributes
        auto to_row = [&](uint32_t c, uint32_t prf) -> uint32_t {
            const uint32_t off = prf & cmask, row = (c << cshift) + off;
            const bool in = c < S && row < n_rows;
            const uint32_t safe = in ? row : (off < n_rows ? off : 0);
            return safe | ((in && active && (int32_t)c != skip) ? ROW_CONTRIB : 0u);
        };
        PrfState st;
        prf_rounds<NTAB, NB, 1, 10>(T, R, g, (uint32_t)gl, st);  // prologue: rows of the first group
        uint32_t row_next = to_row(gl, st.s0);
        for (uint32_t c0 = 0; c0 < S; c0 += G) {
            const uint32_t row_cur = row_next, c_next = c0 + G + gl;
            hg_phase<VT, G, NV, NTAB, NB, U, FULL, 0, NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
            if (NPH > 1) hg_phase<VT, G, NV, NTAB, NB, U, FULL, (NPH > 1 ? 1 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
            if (NPH > 2) {
                hg_phase<VT, G, NV, NTAB, NB, U, FULL, (NPH > 2 ? 2 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
                hg_phase<VT, G, NV, NTAB, NB, U, FULL, (NPH > 2 ? 3 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
            }
            row_next = to_row(c_next, st.s0);
        }
        if (active) {
            VT *o = reinterpret_cast<VT *>(J.out) + i * ev + gl;
#pragma unroll
            for (int k = 0; k < NV; k++)
                if ((uint32_t)(k * G + gl) < ev) o[k * G] = par[k];
            VT z;
            vzero(z);
            for (uint32_t col = NV * G + gl; col < ev; col += G) o[col - gl] = z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A6/A8 server answer: one CTA per sub-query
// ---------------------------------------------------------------------------------------------
constexpr int ANS_THREADS = 256;
struct AnswerParams {
    const void *db;
    const uint64_t *row0, *n_rows;
    const uint32_t *chunk_size, *set_size, *offsets;
    uint64_t offsets_stride;
    uint64_t *out;
    uint32_t ev, evx;
    uint32_t split;   // CTAs per sub-query (a thread-block cluster when > 1)
};

// CW lanes cover the columns of one row; ANS_THREADS/CW rows are fetched concurrently.
// A call with few sub-queries (one search step of one client: 96) leaves SMs idle and is bound by the depth of one CTA's
// row pipeline, so `split` CTAs -- a thread-block cluster -- share a sub-query: each XORs every split-th group of rows,
// and rank 0 folds the partial parities of its peers through distributed shared memory.
template <typename VT, int CW, int NVA>
__global__ void __launch_bounds__(ANS_THREADS) answer_kernel(const AnswerParams P) {
    extern __shared__ uint32_t smem_u32[];
    const uint32_t R = P.split, q = blockIdx.x / R, rank = blockIdx.x % R, t = threadIdx.x;
    const uint32_t S = P.set_size[q], C = P.chunk_size[q];
    const uint64_t n_rows = P.n_rows[q];
    const VT *base = reinterpret_cast<const VT *>(P.db) + P.row0[q] * P.ev;
    uint32_t *s_off = smem_u32;                                             // [S]
    VT *s_red = reinterpret_cast<VT *>(smem_u32 + ((S + 3) & ~3u));         // [groups][CW*NVA]
    VT *s_part = s_red + (ANS_THREADS / CW) * (CW * NVA);                   // [CW*NVA] this CTA's partial parity (split > 1)
    for (uint32_t c = t; c < S; c += ANS_THREADS) s_off[c] = P.offsets[q * P.offsets_stride + c];
    __syncthreads();
    constexpr int GROUPS = ANS_THREADS / CW;
    const uint32_t col = t % CW, grp = t / CW;
    // column blocks of CW*NVA vectors (a single block for rows up to CW*NVA*sizeof(VT) bytes)
    for (uint32_t cb = 0; cb < P.ev; cb += CW * NVA) {
        VT acc[NVA];
#pragma unroll
        for (int k = 0; k < NVA; k++) vzero(acc[k]);
#pragma unroll 4
        for (uint32_t c = grp + rank * GROUPS; c < S; c += GROUPS * R) {
            const uint64_t idx = (uint64_t)s_off[c] + (uint64_t)c * C;
            const bool ok = idx < n_rows;
            const VT *rp = base + (ok ? idx : 0) * P.ev + cb + col;
#pragma unroll
            for (int k = 0; k < NVA; k++) vxor(acc[k], ldg_row(rp + k * CW, ok && cb + k * CW + col < P.evx));
        }
#pragma unroll
        for (int k = 0; k < NVA; k++) s_red[grp * (CW * NVA) + k * CW + col] = acc[k];
        __syncthreads();
        for (uint32_t v = t; v < CW * NVA && cb + v < P.ev; v += ANS_THREADS) {
            VT r;
            vzero(r);
            for (int gidx = 0; gidx < GROUPS; gidx++) vxor(r, s_red[gidx * (CW * NVA) + v]);
            if (R == 1) reinterpret_cast<VT *>(P.out)[(uint64_t)q * P.ev + cb + v] = r;
            else s_part[v] = r;
        }
        if (R > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();                                   // every rank's partial parity is in its shared memory
            if (rank == 0)
                for (uint32_t v = t; v < CW * NVA && cb + v < P.ev; v += ANS_THREADS) {
                    VT r = s_part[v];
                    for (uint32_t pr = 1; pr < R; pr++) vxor(r, cluster.map_shared_rank(s_part, pr)[v]);
                    reinterpret_cast<VT *>(P.out)[(uint64_t)q * P.ev + cb + v] = r;
                }
            cluster.sync();                                   // peers keep their shared memory until rank 0 has read it
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// row gather (replacement values) and xorSlices
// ---------------------------------------------------------------------------------------------
template <typename VT>
__global__ void gather_rows_kernel(const VT *db, uint64_t row0, uint64_t n_rows, uint32_t ev, const uint64_t *idx,
                                   uint64_t n, VT *out) {
    const uint64_t total = n * ev;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = t / ev, col = t % ev, r = idx[i];
        VT v;
        vzero(v);
        if (r < n_rows) v = db[(row0 + r) * ev + col];
        out[t] = v;
    }
}
__global__ void xor_slices_kernel(uint64_t *dst, const uint64_t *src, uint64_t n4) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n4; t += (uint64_t)gridDim.x * blockDim.x)
        dst[t] ^= src[t];
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
// tuning knob (read once): PM_HG_NTAB=1|4 forces the number of T-tables; default by row width (measured on B200:
// four tables win while the PRF dominates, i.e. rows up to 640 B; one table + PRMT rotations wins for 896 B rows)
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
template <typename KERN>
static int launch_hg(KERN kern, int smem, const HintParams &P, int sm, cudaStream_t st) {
    PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint32_t grid = P.n_tiles < (uint32_t)sm ? P.n_tiles : (uint32_t)sm;
    if (P.sync && P.n_tiles > grid) {
        // round barrier needs every CTA resident: cooperative launch (one 512-thread CTA per SM always fits)
        PM_CUDA(cudaMemsetAsync(P.sync, 0, sizeof(unsigned int), st));
        void *args[] = {(void *)&P};
        PM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(P.threads), args, (size_t)smem, st));
    } else {
        HintParams Q = P;
        Q.sync = nullptr;
        kern<<<grid, Q.threads, smem, st>>>(Q);
    }
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
template <typename VT, int G, int NV, int NB>
static int launch_hintgen_t(const HintParams &P, int sm, cudaStream_t st) {
    constexpr int U = (G >= 2) ? 2 : 1;
    constexpr bool kFourTables = sizeof(VT) == 16 && NB == 2;  // instantiated only where it is ever selected
    static const int forced = env_int("PM_HG_NTAB", 0);
    const int ntab = forced ? forced : (NV * G <= 40 ? 4 : 1);
    const bool full = (P.evx == (uint32_t)(NV * G));
    if (kFourTables && ntab == 4) {
        if (full) return launch_hg(hintgen_kernel<VT, G, NV, kFourTables ? 4 : 1, NB, U, true>, aes_tab_words<4>() * 4, P, sm, st);
        return launch_hg(hintgen_kernel<VT, G, NV, kFourTables ? 4 : 1, NB, U, false>, aes_tab_words<4>() * 4, P, sm, st);
    }
    if (full) return launch_hg(hintgen_kernel<VT, G, NV, 1, NB, U, true>, aes_tab_words<1>() * 4, P, sm, st);
    return launch_hg(hintgen_kernel<VT, G, NV, 1, NB, U, false>, aes_tab_words<1>() * 4, P, sm, st);
}
template <typename VT, int G, int NB>
static int launch_hintgen_nv(int nv, const HintParams &P, int sm, cudaStream_t st) {
    switch (nv) {
    case 1: return launch_hintgen_t<VT, G, 1, NB>(P, sm, st);
    case 2: return launch_hintgen_t<VT, G, 2, NB>(P, sm, st);
    case 3: return launch_hintgen_t<VT, G, 3, NB>(P, sm, st);
    case 4: return launch_hintgen_t<VT, G, 4, NB>(P, sm, st);
    case 5: return launch_hintgen_t<VT, G, 5, NB>(P, sm, st);
    case 6: return launch_hintgen_t<VT, G, 6, NB>(P, sm, st);
    case 7: return launch_hintgen_t<VT, G, 7, NB>(P, sm, st);
    case 8: return launch_hintgen_t<VT, G, 8, NB>(P, sm, st);
    }
    return set_error(PM_ERR_UNSUPPORTED, "hintgen: entry too wide (nv=%d)", nv);
}
template <typename VT, int NB>
static int launch_hintgen_g(uint32_t evx, const HintParams &P, int sm, cudaStream_t st, uint32_t *hints_per_tile,
                            bool query_only) {
    const uint32_t x = evx ? evx : 1;
    int G = x >= 8 ? 8 : x >= 4 ? 4 : x >= 2 ? 2 : 1;
    int nv = (int)((x + G - 1) / G);
    *hints_per_tile = (32 / G);   // hints per warp; the caller multiplies by the warps of the CTA width it picks
    if (query_only) return nv <= 8 ? PM_OK : set_error(PM_ERR_UNSUPPORTED, "hintgen: entry_u64 too large");
    switch (G) {
    case 8: return launch_hintgen_nv<VT, 8, NB>(nv, P, sm, st);
    case 4: return nv == 1 ? launch_hintgen_t<VT, 4, 1, NB>(P, sm, st) : launch_hintgen_t<VT, 4, 2, NB>(P, sm, st);
    case 2: return nv == 1 ? launch_hintgen_t<VT, 2, 1, NB>(P, sm, st) : launch_hintgen_t<VT, 2, 2, NB>(P, sm, st);
    default: return launch_hintgen_t<VT, 1, 1, NB>(P, sm, st);
    }
}

// Enqueue hint generation for `jobs` (all pointers inside are device pointers) on `st`.
int hintgen_enqueue(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, cudaStream_t st) {
    const uint64_t E = db->entry_u64;
    const bool wide = (E % 2 == 0);  // 16-byte vectors need 16-byte aligned rows
    const uint32_t ev = (uint32_t)(wide ? E / 2 : E), evx = (uint32_t)(wide ? (E & ~3ull) / 2 : (E & ~3ull));
    bool need4 = false;
    for (uint64_t a = 0; a < n_jobs; a++) {
        const pm_hint_job &J = jobs[a];
        if (J.chunk_size == 0 || (J.chunk_size & (J.chunk_size - 1)))
            return set_error(PM_ERR_ARG, "hintgen: chunk_size %llu is not a power of two", (unsigned long long)J.chunk_size);
        if (J.row0 + J.n_rows > db->n_rows) return set_error(PM_ERR_ARG, "hintgen: job %llu exceeds the table", (unsigned long long)a);
        if (J.n_rows >= 0x7fffffffull || J.set_size >= 0x7fffffffull || J.chunk_size > 0x80000000ull ||
            J.chunk_size * J.set_size > 0xffffffffull)
            return set_error(PM_ERR_UNSUPPORTED, "hintgen: instance too large for 32-bit row offsets");
        if (J.n_hints && !J.parity_out) return set_error(PM_ERR_ARG, "hintgen: parity_out is null");
        if (J.chunk_size > 65536) need4 = true;
    }
    uint64_t a = 0;
    while (a < n_jobs) {
        HintParams P;
        memset(&P, 0, sizeof(P));
        P.db = db->d_rows;
        // The round barrier pays when a sub-PIR's slice does not fit in L2 (measured: 896 B x 200 k rows = 179 MB,
        // -12 % time); for slices that stay L2-resident anyway (SIFT-shaped: 40 MB) it only costs.  PM_HG_SYNC=0|1 forces.
        static const int force_sync = env_int("PM_HG_SYNC", -1);
        uint64_t max_slice = 0;
        for (uint64_t b = a; b < n_jobs && b < a + HG_MAX_JOBS; b++) max_slice = std::max<uint64_t>(max_slice, jobs[b].n_rows * E * 8);
        const bool use_sync = force_sync >= 0 ? force_sync != 0 : max_slice > (64ull << 20);
        P.sync = use_sync ? sync_counter(db) : nullptr;
        P.ev = ev;
        P.evx = evx;
        uint32_t hpw = 0;   // hints per warp
        int rc = wide ? launch_hintgen_g<uint4, 2>(evx, P, 0, st, &hpw, true) : launch_hintgen_g<uint2, 2>(evx, P, 0, st, &hpw, true);
        if (rc != PM_OK) return rc;
        // CTA width.  Every lane group keeps one hint for a whole sweep, so a tile costs one sweep whatever it holds and
        // the CTAs run ceil(tiles / SMs) rounds.  A round gets cheaper with fewer warps, but less than proportionally
        // (measured on B200, MS-MARCO rows: 16 -> 14 warps = -5 % per round, i.e. round time ~ warps + 24), so a
        // narrower CTA only pays when it saves nothing in rounds: 1/8 of the hints (8 GPUs) is 5.2 -> 6 rounds at 16
        // warps and 5.95 -> 6 at 14 (-4.3 %); at 1, 2 and 4 GPUs 16 warps stay best.  PM_HG_WARPS forces a width.
        static const int force_warps = env_int("PM_HG_WARPS", 0);
        uint32_t warps = HG_THREADS / 32;
        {
            uint64_t best = ~0ull;
            for (uint32_t w = 12; w <= (uint32_t)HG_MAX_THREADS / 32; w++) {
                uint64_t t = 0, nj2 = 0;
                for (uint64_t b = a; b < n_jobs && nj2 < HG_MAX_JOBS; b++) {
                    if (jobs[b].n_hints == 0 || jobs[b].n_rows == 0) continue;
                    nj2++;
                    t += (jobs[b].n_hints + (uint64_t)w * hpw - 1) / ((uint64_t)w * hpw);
                }
                const uint64_t g = std::max<uint64_t>(1, std::min<uint64_t>(t, (uint64_t)db->sm_count));
                const uint64_t cost = ((t + g - 1) / g) * (w + 24);
                if (cost <= best) { best = cost; warps = w; }
            }
            if (force_warps >= 1 && force_warps <= HG_MAX_THREADS / 32) warps = (uint32_t)force_warps;
        }
        const uint32_t hpt = hpw * warps;
        P.threads = warps * 32;
        uint32_t tiles = 0;
        uint32_t nj = 0;
        for (; a < n_jobs && nj < HG_MAX_JOBS; a++) {
            const pm_hint_job &J = jobs[a];
            if (J.n_hints == 0) continue;
            if (J.n_rows == 0) {  // an empty instance has all-zero parities
                PM_CUDA(cudaMemsetAsync(J.parity_out, 0, J.n_hints * E * 8, st));
                continue;
            }
            HintJobDev &D = P.jobs[nj++];
            memcpy(D.rk, J.rk, sizeof(D.rk));
            D.row0 = J.row0; D.n_rows = J.n_rows;
            D.hint_begin = J.hint_begin; D.n_hints = J.n_hints; D.n_primary = J.n_primary; D.backup_group = J.backup_group;
            D.tags = J.tags; D.skip = J.skip_chunk; D.out = J.parity_out;
            D.chunk_mask = (uint32_t)(J.chunk_size - 1);
            D.chunk_shift = (uint32_t)__builtin_ctzll(J.chunk_size);
            D.set_size = (uint32_t)J.set_size;
            D.tile_begin = tiles;
            tiles += (uint32_t)((J.n_hints + hpt - 1) / hpt);
        }
        if (nj == 0) break;
        P.n_jobs = nj;
        P.n_tiles = tiles;
        uint32_t unused = 0;
        if (wide) rc = need4 ? launch_hintgen_g<uint4, 4>(evx, P, db->sm_count, st, &unused, false)
                             : launch_hintgen_g<uint4, 2>(evx, P, db->sm_count, st, &unused, false);
        else if (need4) rc = set_error(PM_ERR_UNSUPPORTED, "hintgen: odd entry_u64 with chunk_size > 65536 is not built");
        else rc = launch_hintgen_g<uint2, 2>(evx, P, db->sm_count, st, &unused, false);
        if (rc != PM_OK) return rc;
    }
    return PM_OK;
}

template <typename VT>
static int launch_answer_t(const AnswerParams &P, uint64_t q, uint32_t max_set, cudaStream_t st) {
    const uint32_t x = P.evx ? P.evx : 1;
    const size_t off_words = (max_set + 3) & ~3u;
#define PM_ANS(CW, NVA)                                                                                  \
    do {                                                                                                 \
        size_t smem = off_words * 4 + (size_t)ANS_THREADS * NVA * sizeof(VT) + (size_t)CW * NVA * sizeof(VT); \
        auto kern = answer_kernel<VT, CW, NVA>;                                                          \
        PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        if (P.split > 1) {                                                                               \
            cudaLaunchConfig_t cfg = {};                                                                 \
            cfg.gridDim = dim3((unsigned)(q * P.split));                                                 \
            cfg.blockDim = dim3(ANS_THREADS);                                                            \
            cfg.dynamicSmemBytes = smem;                                                                 \
            cfg.stream = st;                                                                             \
            cudaLaunchAttribute attr[1];                                                                 \
            attr[0].id = cudaLaunchAttributeClusterDimension;                                            \
            attr[0].val.clusterDim.x = P.split;                                                          \
            attr[0].val.clusterDim.y = 1;                                                                \
            attr[0].val.clusterDim.z = 1;                                                                \
            cfg.attrs = attr;                                                                            \
            cfg.numAttrs = 1;                                                                            \
            PM_CUDA(cudaLaunchKernelEx(&cfg, kern, P));                                                  \
        } else {                                                                                         \
            kern<<<(unsigned)q, ANS_THREADS, smem, st>>>(P);                                             \
        }                                                                                                \
    } while (0)
    if (x <= 1) PM_ANS(1, 1);
    else if (x <= 2) PM_ANS(2, 1);
    else if (x <= 4) PM_ANS(4, 1);
    else if (x <= 8) PM_ANS(8, 1);
    else if (x <= 16) PM_ANS(16, 1);
    else if (x <= 32) PM_ANS(32, 1);
    else if (x <= 64) PM_ANS(32, 2);
    else PM_ANS(32, 4);
#undef PM_ANS
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}

int answer_enqueue(pm_db *db, const uint64_t *row0, const uint64_t *n_rows, const uint32_t *chunk_size,
                   const uint32_t *set_size, const uint32_t *offsets, uint64_t stride, uint64_t q, uint32_t max_set,
                   uint64_t *out, cudaStream_t st) {
    if (q == 0) return PM_OK;
    if (q > 0x7fffffffull) return set_error(PM_ERR_UNSUPPORTED, "answer: too many sub-queries in one call");
    if ((size_t)max_set * 4 > 128 * 1024) return set_error(PM_ERR_UNSUPPORTED, "answer: set_size too large");
    const uint64_t E = db->entry_u64;
    const bool wide = (E % 2 == 0);
    AnswerParams P;
    P.db = db->d_rows; P.row0 = row0; P.n_rows = n_rows; P.chunk_size = chunk_size; P.set_size = set_size;
    P.offsets = offsets; P.offsets_stride = stride; P.out = out;
    P.ev = (uint32_t)(wide ? E / 2 : E);
    P.evx = (uint32_t)(wide ? (E & ~3ull) / 2 : (E & ~3ull));
    // few sub-queries: a cluster of up to 4 CTAs per sub-query so that about three CTAs per SM are at work (measured at
    // 96 sub-queries, us: 1 CTA 24.9 | 2 21.2 | 3 18.4 | 4 16.8 | 6 16.5 | 8 20.0).  PM_ANS_SPLIT forces.
    static const int force_split = env_int("PM_ANS_SPLIT", 0);
    const uint64_t target = 3ull * (uint64_t)db->sm_count;
    uint32_t split = (uint32_t)std::min<uint64_t>(4, std::max<uint64_t>(1, (target + q / 2) / q));
    if (max_set < 64) split = 1;   // nothing to share
    if (force_split >= 1 && force_split <= 8) split = (uint32_t)force_split;
    P.split = split;
    return wide ? launch_answer_t<uint4>(P, q, max_set, st) : launch_answer_t<uint2>(P, q, max_set, st);
}

int gather_enqueue(pm_db *db, uint64_t row0, uint64_t n_rows, const uint64_t *idx, uint64_t n, uint64_t *out,
                   cudaStream_t st) {
    if (n == 0) return PM_OK;
    const uint64_t E = db->entry_u64;
    const uint64_t total = n * ((E % 2 == 0) ? E / 2 : E);
    unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (E % 2 == 0)
        gather_rows_kernel<uint4><<<grid, 256, 0, st>>>((const uint4 *)db->d_rows, row0, n_rows, (uint32_t)(E / 2), idx, n, (uint4 *)out);
    else
        gather_rows_kernel<uint2><<<grid, 256, 0, st>>>((const uint2 *)db->d_rows, row0, n_rows, (uint32_t)E, idx, n, (uint2 *)out);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}

}  // namespace pm

using namespace pm;

// =============================================================================================
// C-ABI
// =============================================================================================
PM_EXPORT int pm_expand_key_batch(const uint8_t *keys, uint64_t n, uint32_t *rk) {
    if (n && (!keys || !rk)) return set_error(PM_ERR_ARG, "pm_expand_key: null pointer");
    if (n == 0) return PM_OK;
    if (n > 1u << 24) return set_error(PM_ERR_UNSUPPORTED, "pm_expand_key_batch: too many keys");
    int rc = ensure_device(-1);
    if (rc) return rc;
    void *w = nullptr;
    cudaStream_t st;
    std::unique_lock<std::mutex> lock;
    if ((rc = dev_work(-1, n * (16 + 176), &w, &st, &lock))) return rc;
    uint32_t *d_key = (uint32_t *)w, *d_rk = d_key + 4 * n;
    PM_CUDA(cudaMemcpyAsync(d_key, keys, n * 16, cudaMemcpyHostToDevice, st));
    expand_key_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_key, d_rk, (uint32_t)n);
    PM_CHECK_LAUNCH();
    count_launch();
    PM_CUDA(cudaMemcpyAsync(rk, d_rk, n * 176, cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaStreamSynchronize(st));
    return PM_OK;
}
PM_EXPORT int pm_expand_key(const uint8_t key[16], uint32_t rk[44]) { return pm_expand_key_batch(key, 1, rk); }

PM_EXPORT int pm_prf_batch(const uint32_t rk[44], const uint64_t *tags, const uint64_t *xs, uint64_t n, uint64_t *out) {
    if (!rk || (n && (!tags || !xs || !out))) return set_error(PM_ERR_ARG, "pm_prf_batch: null pointer");
    if (n == 0) return PM_OK;
    int rc = ensure_device(-1);
    if (rc) return rc;
    uint64_t *d = nullptr;
    PM_CUDA(cudaMalloc(&d, 3 * n * 8));
    cudaError_t e = cudaMemcpy(d, tags, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, xs, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        RkArray a;
        memcpy(a.w, rk, sizeof(a.w));
        unsigned grid = (unsigned)((n + 255) / 256 < 148 * 4 ? (n + 255) / 256 : 148 * 4);
        prf_batch_kernel<<<grid, 256, aes_tab_words<1>() * 4>>>(a, d, d + n, n, d + 2 * n);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d + 2 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(PM_ERR_CUDA, "pm_prf_batch: %s", cudaGetErrorString(e));
    return PM_OK;
}

PM_EXPORT int pm_xor_slices(uint64_t *dst, const uint64_t *src, uint64_t len_src) {
    const uint64_t n4 = len_src & ~3ull;  // the reference asm processes len(src)/4 blocks of 4 words
    if (n4 == 0) return PM_OK;
    if (!dst || !src) return set_error(PM_ERR_ARG, "pm_xor_slices: null pointer");
    int rc = ensure_device(-1);
    if (rc) return rc;
    uint64_t *d = nullptr;
    PM_CUDA(cudaMalloc(&d, 2 * n4 * 8));
    cudaError_t e = cudaMemcpy(d, dst, n4 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n4, src, n4 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        unsigned grid = (unsigned)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
        xor_slices_kernel<<<grid, 256>>>(d, d + n4, n4);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(dst, d, n4 * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(PM_ERR_CUDA, "pm_xor_slices: %s", cudaGetErrorString(e));
    return PM_OK;
}

PM_EXPORT int pm_hintgen_dev(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, void *stream) {
    if (!db || (n_jobs && !jobs)) return set_error(PM_ERR_ARG, "pm_hintgen_dev: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    return hintgen_enqueue(db, jobs, n_jobs, stream ? (cudaStream_t)stream : db->stream);
}

// Host-buffer variant: tags/skip uploaded, parities downloaded.  Jobs are issued in a few launch
// groups so the D2H of one group's parities overlaps the next group's kernel.
PM_EXPORT int pm_hintgen(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs) {
    if (!db || (n_jobs && !jobs)) return set_error(PM_ERR_ARG, "pm_hintgen: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint64_t E = db->entry_u64;
    uint64_t total_hints = 0, explicit_tags = 0, explicit_skip = 0;
    for (uint64_t a = 0; a < n_jobs; a++) {
        total_hints += jobs[a].n_hints;
        if (jobs[a].tags) explicit_tags += jobs[a].n_hints;
        if (jobs[a].skip_chunk) explicit_skip += jobs[a].n_hints;
    }
    if (total_hints == 0) return PM_OK;
    void *d_out = nullptr, *d_tags = nullptr, *d_skip = nullptr;
    if ((rc = scratch(db, 0, total_hints * E * 8, &d_out))) return rc;
    if (explicit_tags && (rc = scratch(db, 1, explicit_tags * 8, &d_tags))) return rc;
    if (explicit_skip && (rc = scratch(db, 2, explicit_skip * 4, &d_skip))) return rc;

    std::vector<pm_hint_job> dj(jobs, jobs + n_jobs);
    uint64_t ho = 0, to = 0, so = 0;
    for (uint64_t a = 0; a < n_jobs; a++) {
        dj[a].parity_out = (uint64_t *)d_out + ho * E;
        ho += jobs[a].n_hints;
        if (jobs[a].n_hints && !jobs[a].parity_out) return set_error(PM_ERR_ARG, "pm_hintgen: parity_out is null");
        if (jobs[a].tags) {
            dj[a].tags = (uint64_t *)d_tags + to;
            PM_CUDA(cudaMemcpyAsync((void *)dj[a].tags, jobs[a].tags, jobs[a].n_hints * 8, cudaMemcpyHostToDevice, db->stream));
            to += jobs[a].n_hints;
        }
        if (jobs[a].skip_chunk) {
            dj[a].skip_chunk = (int32_t *)d_skip + so;
            PM_CUDA(cudaMemcpyAsync((void *)dj[a].skip_chunk, jobs[a].skip_chunk, jobs[a].n_hints * 4, cudaMemcpyHostToDevice, db->stream));
            so += jobs[a].n_hints;
        }
    }
    const uint64_t groups = n_jobs < 4 ? n_jobs : 4;
    for (uint64_t gi = 0; gi < groups; gi++) {
        const uint64_t a0 = n_jobs * gi / groups, a1 = n_jobs * (gi + 1) / groups;
        if ((rc = hintgen_enqueue(db, dj.data() + a0, a1 - a0, db->stream))) return rc;
        PM_CUDA(cudaEventRecord(db->ev[gi & 3], db->stream));
        PM_CUDA(cudaStreamWaitEvent(db->copy_stream, db->ev[gi & 3], 0));
        for (uint64_t a = a0; a < a1; a++)
            if (jobs[a].n_hints)
                PM_CUDA(cudaMemcpyAsync(jobs[a].parity_out, dj[a].parity_out, jobs[a].n_hints * E * 8, cudaMemcpyDeviceToHost, db->copy_stream));
    }
    PM_CUDA(cudaStreamSynchronize(db->copy_stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

PM_EXPORT int pm_gather_rows(pm_db *db, uint64_t row0, uint64_t n_rows, const uint64_t *idx, uint64_t n, uint64_t *out) {
    if (!db || (n && (!idx || !out))) return set_error(PM_ERR_ARG, "pm_gather_rows: null pointer");
    if (row0 + n_rows > db->n_rows) return set_error(PM_ERR_ARG, "pm_gather_rows: slice exceeds the table");
    if (n == 0) return PM_OK;
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint64_t E = db->entry_u64;
    void *d_idx = nullptr, *d_out = nullptr;
    if ((rc = scratch(db, 1, n * 8, &d_idx))) return rc;
    if ((rc = scratch(db, 0, n * E * 8, &d_out))) return rc;
    PM_CUDA(cudaMemcpyAsync(d_idx, idx, n * 8, cudaMemcpyHostToDevice, db->stream));
    if ((rc = gather_enqueue(db, row0, n_rows, (const uint64_t *)d_idx, n, (uint64_t *)d_out, db->stream))) return rc;
    PM_CUDA(cudaMemcpyAsync(out, d_out, n * E * 8, cudaMemcpyDeviceToHost, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

static int answer_check(pm_db *db, const void *a, const void *b, const void *c, const void *d, const void *e, const void *f, uint64_t q) {
    if (!db) return set_error(PM_ERR_ARG, "pm_answer_batch: null handle");
    if (q && (!a || !b || !c || !d || !e || !f)) return set_error(PM_ERR_ARG, "pm_answer_batch: null pointer");
    return PM_OK;
}

PM_EXPORT int pm_answer_batch_dev(pm_db *db, const uint64_t *row0, const uint64_t *n_rows, const uint32_t *chunk_size,
                                  const uint32_t *set_size, const uint32_t *offsets, uint64_t offsets_stride, uint64_t q,
                                  uint64_t *out, void *stream) {
    int rc = answer_check(db, row0, n_rows, chunk_size, set_size, offsets, out, q);
    if (rc) return rc;
    if ((rc = ensure_device(db->device))) return rc;
    // device-side descriptors cannot be validated here; offsets_stride bounds the per-query set size
    return answer_enqueue(db, row0, n_rows, chunk_size, set_size, offsets, offsets_stride, q, (uint32_t)offsets_stride, out,
                          stream ? (cudaStream_t)stream : db->stream);
}

PM_EXPORT int pm_answer_batch(pm_db *db, const uint64_t *row0, const uint64_t *n_rows, const uint32_t *chunk_size,
                              const uint32_t *set_size, const uint32_t *offsets, uint64_t offsets_stride, uint64_t q,
                              uint64_t *out) {
    int rc = answer_check(db, row0, n_rows, chunk_size, set_size, offsets, out, q);
    if (rc) return rc;
    if (q == 0) return PM_OK;
    uint32_t max_set = 0;
    for (uint64_t i = 0; i < q; i++) {
        if (row0[i] + n_rows[i] > db->n_rows) return set_error(PM_ERR_ARG, "pm_answer_batch: sub-query %llu exceeds the table", (unsigned long long)i);
        if (set_size[i] > offsets_stride) return set_error(PM_ERR_ARG, "pm_answer_batch: set_size > offsets_stride");
        if (set_size[i] > max_set) max_set = set_size[i];
    }
    if ((rc = ensure_device(db->device))) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint64_t E = db->entry_u64;
    // one staging block: row0[q] n_rows[q] (u64) | chunk[q] set[q] (u32) | offsets[q*stride] (u32)
    const size_t bytes_desc = q * (8 + 8 + 4 + 4), bytes_off = q * offsets_stride * 4;
    void *d_in = nullptr, *d_out = nullptr;
    if ((rc = scratch(db, 1, bytes_desc + bytes_off, &d_in))) return rc;
    if ((rc = scratch(db, 0, q * E * 8, &d_out))) return rc;
    uint64_t *d_row0 = (uint64_t *)d_in, *d_nrows = d_row0 + q;
    uint32_t *d_chunk = (uint32_t *)(d_nrows + q), *d_set = d_chunk + q, *d_off = d_set + q;
    PM_CUDA(cudaMemcpyAsync(d_row0, row0, q * 8, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_nrows, n_rows, q * 8, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_chunk, chunk_size, q * 4, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_set, set_size, q * 4, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_off, offsets, bytes_off, cudaMemcpyHostToDevice, db->stream));
    if ((rc = answer_enqueue(db, d_row0, d_nrows, d_chunk, d_set, d_off, offsets_stride, q, max_set, (uint64_t *)d_out, db->stream))) return rc;
    PM_CUDA(cudaMemcpyAsync(out, d_out, q * E * 8, cudaMemcpyDeviceToHost, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

#include "pm_l2.cuh"
#include "pm_client.cuh"
