// A11 for many queries at once: the uint32 wrapping inner-product scan (graphann/l2_distance_amd64.s:39-68,
// graphann_test.go:268-273) as an int8 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators
// kept in TMEM across the whole row sweep, operands staged by TMA) -- the one place on this path where tensor cores are the right tool
// (BASELINE.json north_star; SURVEY.md 7.5: 6.15e11 32-bit MACs at Q = 1000 are integer-pipe bound).
//
// Formulation.  A u32 x u32 product mod 2^32 is  sum_{a+b<=3} A_a * B_b * 2^(8(a+b))  over the byte limbs, i.e. ten
// u8 x u8 -> s32 GEMMs grouped by shift s = a + b into four accumulators D_s, and IP(i,t) = sum_s D_s[i][t] << 8s.
// Every partial sum is <= 4*dim*255^2 < 2^31 for dim <= 8192: exact.  Only the checksum sum_i IP(i,t) leaves the
// kernel: the epilogue folds the four accumulators and sums over rows.
//
// Operands.  The row table is consumed AS IT LIES IN MEMORY: TMA drops 128 rows x 128 B (32 elements per row) into
// shared memory; each sorter thread reads its row, sorts the bytes by limb in registers -- [limb 0 of the 32 elements |
// limb 1 | limb 2 | limb 3], 32 B each -- and stores them with tcgen05.st into its TMEM lane: the MMA takes A FROM
// TMEM (lane = row, four K bytes per 32-bit column, 8 columns per K = 32 slice).  The query operand is written
// limb-sorted the same way by a small kernel ([Q][4*dim] bytes, one copy) and staged by TMA in shared memory.  Limb
// pair (a, b) is then the MMA of A slice a with B slice b of the same 128-byte chunk: 10 MMAs per chunk.
//
// Why A lives in TMEM: with both operands in shared memory an M128 N128 K32 MMA reads 8 KB for 64 clk of math, i.e.
// the whole 128 B/clk of the SM's shared memory, with the TMA fills and the limb sort on top.  With A in TMEM the
// shared-memory traffic per chunk drops from 144 KB to 67 KB, and a stage of the ring is dead as soon as its sorters
// hold it in registers.  TMEM budget: 4 accumulators x 112 columns + 2 A buffers x 32 columns = 512, hence the query
// tile of 112.
//
// Work split: CTA b takes query tile b % q_slots and row tiles b / q_slots (+ groups, ...).  The q_slots CTAs of a
// group sweep the same row tiles at the same pace, so a row tile comes from HBM once (ncu: 2.5 GB for a 2.46 GB table,
// L2 hit rate 86 %) and the CTA's query tile stays resident in shared memory for the whole sweep.
//
// Kernel: one persistent CTA per SM, warp-specialised (11 warps).  One lane of the TMA warp is the producer of an
// 8-stage mbarrier ring of 16 KB A tiles (A + B tiles when 4*dim bytes of queries do not fit shared memory); two groups
// of four warps sort the limbs of alternate A tiles as they land (thread = row = TMEM lane), release the stage, store
// the limbs to TMEM and signal `sorted`; two MMA warps take alternate chunks and one lane of each issues the
// tcgen05.mma (M128 N112 K32, kind::i8, B K-major with the 128-byte swizzle the tensor map writes) and commits `empty`,
// which releases the A buffer in TMEM to its sorters.  Warps 0-3 are also the epilogue (warp w reads TMEM lanes
// 32w..32w+31 = rows of the tile) once per query tile.  Every mbarrier wait is bounded: on a timeout the kernel raises
// an error flag and drains instead of hanging.
//
// What the measurements said on the way (scripts/trace_ipgemm.py, profiles/): (1) tcgen05.mma issued from a divergent
// `if (lane == 0)` branch compiles to an ELECT / BRA.U.ANY loop per instruction and the issuer became the bottleneck
// (1240 clk per chunk for 640 clk of math) -- the issuing warps run convergently and elect one lane instead; (2) the
// tensor pipe's queue is shallow, so whatever one issuer does between chunks idles the pipe -- hence two issuers;
// (3) TMA latency under load is ~3000 clk, so the ring must be released by the sorters, not by the MMAs.
// Q = 1000 over 3 201 821 x 192: 3.31 ms = 1.9e15 int8 MAC/s (84 % of the dense int8 peak), integer-pipe kernel 75 ms.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "pm_common.cuh"

namespace pm {

constexpr int G_SORT_GROUPS = 2, G_SORTERS = 128;   // sorter group g = warps 4g..4g+3 takes chunks j = g (mod G_SORT_GROUPS)
constexpr int G_WARP_TMA = 4 * G_SORT_GROUPS, G_WARP_MMA = G_WARP_TMA + 1, G_ISSUERS = 2, G_THREADS = 32 * (G_WARP_MMA + G_ISSUERS);
#ifndef PM_G_TILE_N
#define PM_G_TILE_N 112   // measured at Q = 1000: N = 112 with two A buffers 3.31 ms, N = 96 with four 3.76 ms (11 instead of 9
#define PM_G_ABUFS 2     // readers per row tile: the L2 -> SM traffic, ~7 TB/s, becomes the limit)
#endif
constexpr int G_TILE_M = 128, G_TILE_N = PM_G_TILE_N, G_KCHUNK = 128, G_UMMA_K = 32, G_MAX_STAGES = 8, G_ABUFS = PM_G_ABUFS;
constexpr uint32_t G_TILE_BYTES = G_TILE_M * G_KCHUNK;           // A tile, 16 KB
constexpr uint32_t G_BTILE_BYTES = G_TILE_N * G_KCHUNK;          // B tile, 12 KB
constexpr uint32_t G_STAGE_BYTES = 2 * G_TILE_BYTES;             // A + B (B padded so that every tile stays 1024-byte aligned)
constexpr uint32_t G_TMEM_COLS = 512;
constexpr uint32_t G_TMEM_A = 4 * G_TILE_N;                      // columns 384..511: four limb-sorted A tiles of 4 x 8 columns
static_assert(G_TMEM_A + G_ABUFS * 32 <= G_TMEM_COLS, "accumulators + A buffers must fit TMEM");

// ---- B'_s operand ------------------------------------------------------------------------------------------
// bmat[t][128*g + 32*b + e] = byte b of q_t[32*g + e]   (rows t >= nq are zero)
__global__ void ipgemm_build_b_kernel(const uint32_t *queries, uint32_t nq, uint32_t q_pad, uint32_t dim, uint8_t *bmat) {
    const uint64_t kbytes = (uint64_t)dim * 4, total = (uint64_t)q_pad * kbytes;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t kp = (uint32_t)(i % kbytes), t = (uint32_t)(i / kbytes);
        const uint32_t g = kp >> 7, b = (kp >> 5) & 3, e = kp & 31;
        bmat[i] = t < nq ? (uint8_t)(queries[(uint64_t)t * dim + 32 * g + e] >> (8 * b)) : (uint8_t)0;
    }
}

// ---- PTX helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: false on timeout (~2 s), so a programming error can never hang the GPU
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try(addr, parity))
        if (clock64() - t0 > 4000000000ll) return false;
    return true;
}
// one lane of a converged warp (the warp stays convergent, so the compiler keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>):
// start address >> 4 | LBO = 1 | SBO = 64 (1024 B) | version = 1 | layout = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fff) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// UMMA::InstrDescriptor for kind::i8: D = s32 (bits 4-5 = 2), A = B = uint8 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_u8(int m, int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// A operand from TMEM (lane = row, 4 bytes of K per 32-bit column), B operand from shared memory
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                   "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                   "r"(r[30]), "r"(r[31]) : "memory");
}

struct GemmParams {
    uint64_t n_rows;
    uint32_t n_row_tiles, n_q_tiles, q_pad, k_chunks;
    uint32_t q_slots, groups;       // CTA b works on query tiles b % q_slots (+ q_slots, ...) and row tiles b / q_slots (+ groups, ...)
    uint32_t n_stages, stage_bytes; // A ring (16 KB stages) when the B tile is resident, A + B ring (32 KB stages) otherwise
    uint32_t b_resident;            // bytes of the resident B tile (k_chunks * G_BTILE_BYTES) or 0
    uint32_t *checksum;             // [q_pad]
    int *error_flag;
    unsigned long long *trace;      // debug build (-DPM_IPGEMM_TRACE): per-chunk clock stamps of CTA 0, else unused
};
#ifdef PM_IPGEMM_TRACE
#define G_TRACE_N 512
#define G_TRACE(role, j) do { if (blockIdx.x == 0 && (j) < G_TRACE_N && (threadIdx.x & 31) == 0) P.trace[(role) * G_TRACE_N + (j)] = clock64(); } while (0)
#else
#define G_TRACE(role, j) do { } while (0)
#endif

// Limb sort of one 128-byte row of the A tile (32 uint32 elements) into registers: out[8a .. 8a+7] = byte a of the 32
// elements, in element order.  The row is stored as eight 16-byte chunks, chunk c at physical position c ^ (row & 7)
// (the 128-byte swizzle TMA wrote).
__device__ __forceinline__ void limb_sort_row(const uint8_t *tile, uint32_t row, uint32_t (&out)[32]) {
    const uint4 *rowp = reinterpret_cast<const uint4 *>(tile + row * 128);
    const uint32_t x = row & 7;
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 v = rowp[c ^ x];
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int e = 4 * m;
            const uint32_t lo = __byte_perm(w[e], w[e + 1], 0x0040 + a * 0x0011);        // (w[e].b_a, w[e+1].b_a, -, -)
            const uint32_t hi = __byte_perm(w[e + 2], w[e + 3], 0x0040 + a * 0x0011);
            out[8 * a + m] = __byte_perm(lo, hi, 0x5410);
        }
}

__global__ void __launch_bounds__(G_THREADS, 1) ipgemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b, const GemmParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem_b = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem = smem_b + P.b_resident;   // the ring
    __shared__ uint64_t bar_full[G_MAX_STAGES], bar_sorted[G_MAX_STAGES], bar_empty[G_MAX_STAGES], bar_a_read[G_MAX_STAGES], bar_accum, bar_tmem_free, bar_b, bar_b_free, bar_first;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform

    if (threadIdx.x == 0) {
        for (int i = 0; i < G_MAX_STAGES; i++) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_sorted[i], G_SORTERS);
            mbar_init(&bar_empty[i], 1);
            mbar_init(&bar_a_read[i], G_SORTERS);
        }
        mbar_init(&bar_accum, G_ISSUERS);
        mbar_init(&bar_tmem_free, G_SORTERS);
        mbar_init(&bar_b, 1);
        mbar_init(&bar_b_free, G_ISSUERS);
        mbar_init(&bar_first, 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 0) {  // one warp allocates all 512 TMEM columns and publishes the base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(G_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem_base;

    // This CTA's work.  CTAs that share b / q_slots sweep the SAME row tiles at the same pace, each against its own
    // query tile, so a row tile comes from HBM once and from L2 for the other q_slots - 1 readers; the query tile of a
    // CTA stays resident in shared memory for the whole sweep when it fits.  Flattened iteration
    // j = ((qi * my_tiles) + r) * k_chunks + kc,  query tile = slot + qi * q_slots,  row tile = rg + r * groups.
    // The accumulators are NOT drained per row tile: D_s keeps accumulating over all row tiles of a query tile.  s32
    // accumulation wraps mod 2^32 and only sum_i D_s[i][t] << 8s mod 2^32 is wanted, so the wrap is harmless, and the
    // TMEM read-out (224 KB) happens once per query tile instead of once per (query tile, row tile).
    const uint32_t slot = blockIdx.x % P.q_slots, rg = blockIdx.x / P.q_slots;
    const uint32_t my_tiles = P.n_row_tiles > rg ? (P.n_row_tiles - rg + P.groups - 1) / P.groups : 0;
    const uint32_t my_qts = P.n_q_tiles > slot ? (P.n_q_tiles - slot + P.q_slots - 1) / P.q_slots : 0;
    const uint32_t per_qt = my_tiles * P.k_chunks, total = per_qt * my_qts;
    const uint32_t n_stages = P.n_stages;
    volatile int *abort_flag = &s_abort;

    if (warp == G_WARP_TMA) {
        // ---- TMA producer: the warp runs the loop convergently, one elected lane issues ----
        uint32_t qi = 0, r = 0, kc = 0, stage = 0, ring = 0;
        for (uint32_t j = 0; j < total; j++) {
            const int32_t q_row = (int32_t)((slot + qi * P.q_slots) * G_TILE_N);
            if (P.b_resident && r == 0 && kc == 0) {   // (re)load this CTA's query tile: all K chunks, one barrier
                if (qi > 0 && !mbar_wait(&bar_b_free, (qi - 1) & 1)) { *abort_flag = 7; break; }
                if (elect_one()) {
                    mbar_expect_tx(&bar_b, P.b_resident);
                    for (uint32_t c = 0; c < P.k_chunks; c++)
                        tma_load_2d(smem_b + c * G_BTILE_BYTES, &map_b, &bar_b, (int32_t)(c * G_KCHUNK), q_row);
                }
                __syncwarp();
            }
            // An A tile is dead once its sorters hold it in registers (it reaches the MMA through TMEM), so with the
            // query tile resident the stage is refilled without waiting for the MMAs -- the ring then covers the TMA
            // latency (~3000 clk under load) instead of TMA + sort + MMA.  A stage that also carries B waits for the MMAs.
            if (j >= n_stages && !mbar_wait(P.b_resident ? &bar_a_read[stage] : &bar_empty[stage], ring ^ 1)) { *abort_flag = 1; break; }
            G_TRACE(0, j);
            if (elect_one()) {
                uint8_t *sa = smem + stage * P.stage_bytes;
                mbar_expect_tx(&bar_full[stage], P.b_resident ? G_TILE_BYTES : G_TILE_BYTES + G_BTILE_BYTES);
                tma_load_2d(sa, &map_a, &bar_full[stage], (int32_t)(kc * G_KCHUNK), (int32_t)((rg + r * P.groups) * G_TILE_M));
                if (!P.b_resident) tma_load_2d(sa + G_TILE_BYTES, &map_b, &bar_full[stage], (int32_t)(kc * G_KCHUNK), q_row);
            }
            __syncwarp();
            if (++kc == P.k_chunks) { kc = 0; if (++r == my_tiles) { r = 0; qi++; } }
            if (++stage == n_stages) { stage = 0; ring ^= 1; }
        }
    } else if (warp >= G_WARP_MMA) {
        // ---- MMA issuers: two warps take alternate chunks of a query tile; each runs its loop convergently and one
        // elected lane issues the 10 tcgen05.mma of a sorted chunk.  Why two: the tensor pipe's queue is shallow (the
        // issue of an MMA blocks until the previous one is nearly done), so whatever the issuer does between two
        // chunks -- barrier round trip (~150 clk per try_wait), fence, descriptor arithmetic, commit -- idles the pipe;
        // with two issuers one prepares while the other issues.  Accumulation is commutative, so the only order that
        // matters is that the accumulate = 0 MMAs of a query tile's first chunk are issued first (bar_first).
        const uint32_t me = warp - G_WARP_MMA;
        const uint32_t idesc = umma_idesc_u8(G_TILE_M, G_TILE_N);
        const uint32_t b_res = P.b_resident, k_chunks = P.k_chunks, stage_bytes = P.stage_bytes;
        const uint32_t ring_addr = smem_u32(smem), bres_addr = smem_u32(smem_b);
        uint32_t rem = 0, kc = 0, qi = 0, stage = 0, ring = 0;
        for (uint32_t j = 0; j < total; j++) {
            const bool last = rem + 1 == per_qt;
            // Both issuers, also one that owns no chunk of this tile: the epilogue of the previous query tile must have
            // read TMEM out -- which also keeps an idle issuer from arriving on bar_accum a whole tile early.
            if (rem == 0 && qi > 0 && !mbar_wait(&bar_tmem_free, (qi - 1) & 1)) { *abort_flag = 4; break; }
            if ((rem & 1u) == me) {
                if (!mbar_wait(&bar_sorted[stage], ring)) { *abort_flag = 2; break; }
                if (!b_res && !mbar_wait(&bar_full[stage], ring)) { *abort_flag = 2; break; }   // streamed B tile: observe the TMA barrier itself
                if (rem == 0) {
                    if (b_res && !mbar_wait(&bar_b, qi & 1)) { *abort_flag = 8; break; }
                } else if (rem == 1) {
                    if (!mbar_wait(&bar_first, qi & 1)) { *abort_flag = 9; break; }   // the other issuer has started this query tile
                }
                G_TRACE(4, j);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_tmem = tmem_base + G_TMEM_A + (j % G_ABUFS) * 32;
                const uint32_t b_addr = b_res ? bres_addr + kc * G_BTILE_BYTES : ring_addr + stage * stage_bytes + G_TILE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (uint32_t a = 0; a < 4; a++)
#pragma unroll
                        for (uint32_t b = 0; a + b < 4; b++)   // limb pair (a, b) accumulates into D_{a+b}; the first pair of a shift is a = 0
                            umma_i8_ts(tmem_base + (a + b) * G_TILE_N, a_tmem + a * 8, umma_desc_sw128(b_addr + b * G_UMMA_K), idesc,
                                       (rem != 0 || a != 0) ? 1u : 0u);
                    umma_commit(&bar_empty[stage]);   // the stage may be refilled and the A buffer rewritten once these MMAs have read them
                    if (rem == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_first)) : "memory");
                }
                __syncwarp();
                G_TRACE(5, j);
            }
            if (last) {   // both issuers: every MMA of mine for this query tile has been issued; the tile's B may be replaced after them
                if (elect_one()) {
                    umma_commit(&bar_accum);
                    if (b_res) umma_commit(&bar_b_free);
                }
                __syncwarp();
                rem = 0;
                qi++;
            } else {
                rem++;
            }
            if (++kc == k_chunks) kc = 0;
            if (++stage == n_stages) { stage = 0; ring ^= 1; }
        }
    } else {
        // ---- limb sorters (thread = row of the A tile = TMEM lane) and epilogue ----
        // group g sorts chunks j = g (mod G_SORT_GROUPS); group 0 is also the epilogue
        const uint32_t group = warp >> 2, row = threadIdx.x & (G_SORTERS - 1);
        const uint32_t lane_addr = tmem_base + (((warp & 3u) * 32u) << 16);
        uint32_t rem = 0, qi = 0, stage = 0, ring = 0;
        uint32_t pstage = 0, pring = 0;   // stage / ring of chunk j - G_ABUFS
        for (uint32_t j = 0; j < total; j++) {
            if (j % G_SORT_GROUPS == group) {
                if (!mbar_wait(&bar_full[stage], ring)) { *abort_flag = 3; break; }
                if ((warp & 3) == 0) G_TRACE(1, j);
                uint32_t limbs[32];
                limb_sort_row(smem + stage * P.stage_bytes, row, limbs);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_a_read[stage])), "r"(limbs[0]), "r"(limbs[31]) : "memory");
                // A buffer j % G_ABUFS in TMEM was last read by the MMAs of chunk j - G_ABUFS
                if (j >= (uint32_t)G_ABUFS) {
                    if (!mbar_wait(&bar_empty[pstage], pring)) { *abort_flag = 6; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                if ((warp & 3) == 0) G_TRACE(2, j);
                tmem_st32(lane_addr + G_TMEM_A + (j % G_ABUFS) * 32, limbs);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_sorted[stage])) : "memory");
                if ((warp & 3) == 0) G_TRACE(3, j);
            }
            if (j >= (uint32_t)G_ABUFS && ++pstage == n_stages) { pstage = 0; pring ^= 1; }
            if (++stage == n_stages) { stage = 0; ring ^= 1; }
            if (++rem != per_qt) continue;
            rem = 0;
            qi++;
            if (group != 0) continue;
            // end of a query tile: fold the four shifted accumulators, sum over the rows, publish
            if (!mbar_wait(&bar_accum, (qi - 1) & 1)) { *abort_flag = 5; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t q0 = (slot + (qi - 1) * P.q_slots) * G_TILE_N;
#pragma unroll 1
            for (int c0 = 0; c0 < G_TILE_N; c0 += 16) {
                uint32_t d0[16], d1[16], d2[16], d3[16];
                tmem_ld16(lane_addr + 0 * G_TILE_N + c0, d0);
                tmem_ld16(lane_addr + 1 * G_TILE_N + c0, d1);
                tmem_ld16(lane_addr + 2 * G_TILE_N + c0, d2);
                tmem_ld16(lane_addr + 3 * G_TILE_N + c0, d3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    uint32_t v = d0[e] + (d1[e] << 8) + (d2[e] << 16) + (d3[e] << 24);
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0 && v) atomicAdd(P.checksum + q0 + c0 + e, v);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_tmem_free)) : "memory");
        }
    }
    __syncthreads();
    if (s_abort && threadIdx.x == 0) atomicExch(P.error_flag, s_abort);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(G_TMEM_COLS) : "memory");
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t row_bytes, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error(PM_ERR_CUDA, "ip gemm: cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {row_bytes, rows}, strides[1] = {row_bytes};
    cuuint32_t box[2] = {G_KCHUNK, box_rows}, estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PM_ERR_CUDA, "ip gemm: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PM_OK;
}

bool ipgemm_applicable(const pm_db *db, uint64_t dim, uint64_t nq, const uint32_t *ip_out) {
    static const int mode = [] { const char *v = getenv("PM_IP_GEMM"); return v && *v ? atoi(v) : 1; }();
    return mode != 0 && ip_out == nullptr && nq >= 64 && dim % 32 == 0 && dim <= 8192 && db->n_rows >= 1 &&
           db->n_rows < (1ull << 31) && ((uintptr_t)db->d_rows % 16) == 0;
}

// checksum[t] = sum_i InnerProduct(row_i, q_t) mod 2^32 for all nq queries; `scratch_dev` must hold
// q_pad*dim*4 bytes (limb-sorted queries) + q_pad*4 (padded checksums) + 16.  Enqueues on st; *err_host is valid after a sync.
int ipgemm_enqueue(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t nq, uint32_t *checksum, void *scratch_dev, cudaStream_t st) {
    const uint32_t q_pad = (uint32_t)((nq + G_TILE_N - 1) / G_TILE_N * G_TILE_N);
    const uint64_t kbytes = dim * 4;
    uint8_t *bmat = (uint8_t *)scratch_dev;
    uint32_t *cs_pad = (uint32_t *)(bmat + (uint64_t)q_pad * kbytes);
    int *err = (int *)(cs_pad + q_pad);
    PM_CUDA(cudaMemsetAsync(cs_pad, 0, (size_t)q_pad * 4 + 16, st));
    ipgemm_build_b_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)q_pad * kbytes + 255) / 256, 148 * 16), 256, 0, st>>>(queries, (uint32_t)nq, q_pad,
                                                                                                                  (uint32_t)dim, bmat);
    PM_CHECK_LAUNCH();
    count_launch();
    CUtensorMap map_a, map_b;
    int rc;
    if ((rc = make_map(&map_a, db->d_rows, db->n_rows, kbytes, G_TILE_M))) return rc;
    if ((rc = make_map(&map_b, bmat, q_pad, kbytes, G_TILE_N))) return rc;
    GemmParams P;
    P.n_rows = db->n_rows;
    P.n_row_tiles = (uint32_t)((db->n_rows + G_TILE_M - 1) / G_TILE_M);
    P.n_q_tiles = q_pad / G_TILE_N;
    P.q_pad = q_pad;
    P.k_chunks = (uint32_t)(kbytes / G_KCHUNK);
    P.checksum = cs_pad;
    P.error_flag = err;
    P.trace = nullptr;
#ifdef PM_IPGEMM_TRACE
    static unsigned long long *d_trace = nullptr;
    if (!d_trace) PM_CUDA(cudaMalloc(&d_trace, 6 * G_TRACE_N * 8));
    PM_CUDA(cudaMemsetAsync(d_trace, 0, 6 * G_TRACE_N * 8, st));
    P.trace = d_trace;
#endif
    // shared memory: the CTA's query tile resident (all K chunks) + a ring of A tiles, or a ring of A + B tiles
    const size_t smem_avail = 227 * 1024 - 2048;
    const size_t b_all = (size_t)P.k_chunks * G_BTILE_BYTES;
    if (b_all + 4 * G_TILE_BYTES <= smem_avail) {
        P.b_resident = (uint32_t)b_all;
        P.stage_bytes = G_TILE_BYTES;
        P.n_stages = (uint32_t)std::min<size_t>(G_MAX_STAGES, (smem_avail - b_all) / G_TILE_BYTES);
    } else {
        P.b_resident = 0;
        P.stage_bytes = G_STAGE_BYTES;
        P.n_stages = 6;
    }
    // CTA grid = q_slots x groups <= SMs.  The table is swept ceil(n_q_tiles / q_slots) times from HBM, so take the most
    // query tiles side by side whose busiest CTA stays within 6 % of the fewest chunks any split achieves.
    const uint32_t sms = (uint32_t)db->sm_count, qs_max = std::min(P.n_q_tiles, sms);
    auto chunks_of = [&](uint32_t qs) {
        const uint32_t g = std::max(1u, std::min(sms / qs, P.n_row_tiles));
        return (uint64_t)((P.n_q_tiles + qs - 1) / qs) * ((P.n_row_tiles + g - 1) / g);
    };
    uint64_t best = ~0ull;
    for (uint32_t qs = 1; qs <= qs_max; qs++) best = std::min(best, chunks_of(qs));
    P.q_slots = 1;
    for (uint32_t qs = 1; qs <= qs_max; qs++)
        if (chunks_of(qs) * 100 <= best * 106) P.q_slots = qs;
    P.groups = std::max(1u, std::min(sms / P.q_slots, P.n_row_tiles));
    const size_t smem = (size_t)P.b_resident + (size_t)P.n_stages * P.stage_bytes + 1024;
    PM_CUDA(cudaFuncSetAttribute(ipgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ipgemm_kernel<<<P.q_slots * P.groups, G_THREADS, smem, st>>>(map_a, map_b, P);
    PM_CHECK_LAUNCH();
    count_launch();
#ifdef PM_IPGEMM_TRACE
    {
        static unsigned long long h[6 * G_TRACE_N];
        PM_CUDA(cudaStreamSynchronize(st));
        PM_CUDA(cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost));
        unsigned long long t0 = h[0];
        fprintf(stderr, "j,tma_issue,full_seen,st_begin,sorted_arrived,mma_sorted_seen,mma_issued\n");
        for (int j = 0; j < G_TRACE_N; j++) {
            fprintf(stderr, "%d", j);
            for (int r = 0; r < 6; r++) fprintf(stderr, ",%lld", h[r * G_TRACE_N + j] ? (long long)(h[r * G_TRACE_N + j] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
    }
#endif
    PM_CUDA(cudaMemcpyAsync(checksum, cs_pad, nq * 4, cudaMemcpyDeviceToDevice, st));
    return PM_OK;
}
size_t ipgemm_scratch_bytes(uint64_t dim, uint64_t nq) {
    const uint64_t q_pad = (nq + G_TILE_N - 1) / G_TILE_N * G_TILE_N;
    return q_pad * dim * 4 + q_pad * 4 + 64;
}
// reads the error flag written by the kernel (after the stream has been synchronised)
int ipgemm_check(void *scratch_dev, uint64_t dim, uint64_t nq) {
    const uint64_t q_pad = (nq + G_TILE_N - 1) / G_TILE_N * G_TILE_N;
    int err = 0;
    PM_CUDA(cudaMemcpy(&err, (uint8_t *)scratch_dev + q_pad * dim * 4 + q_pad * 4, 4, cudaMemcpyDeviceToHost));
    if (err) return set_error(PM_ERR_CUDA, "ip gemm: tensor-core pipeline timed out (code %d)", err);
    return PM_OK;
}

}  // namespace pm
