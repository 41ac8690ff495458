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
// shared memory, and each thread then sorts the bytes of its row by limb in place -- [limb 0 of 32 elements | limb 1 |
// limb 2 | limb 3], 32 B each -- so that every UMMA K = 32 slice of the 128-byte swizzled row is one pure limb plane.
// The query operand is written limb-sorted the same way by a small kernel ([Q][4*dim] bytes, one copy).  Limb pair
// (a, b) is then the MMA of A slice a with B slice b of the same 128-byte chunk: 10 MMAs per chunk.
//
// Kernel: one persistent CTA per SM, 128 threads.  Thread 0 is TMA producer (4 chunks ahead in a 6-stage mbarrier
// ring of 16 KB A + 16 KB B) and tcgen05.mma issuer (single-thread instruction, M128 N128 K32, kind::i8, both operands
// K-major with the 128-byte swizzle the tensor maps write); the four accumulators fill all 512 TMEM columns; all four
// warps sort limbs and are the epilogue (warp w reads TMEM lanes 32w..32w+31 = rows of the tile).  Every mbarrier
// wait is bounded: on a timeout the kernel raises an error flag and drains instead of hanging.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "pm_common.cuh"

namespace pm {

constexpr int G_THREADS = 128;
constexpr int G_TILE_M = 128, G_TILE_N = 128, G_KCHUNK = 128, G_UMMA_K = 32, G_STAGES = 6, G_PREFETCH = 4;
constexpr uint32_t G_TILE_BYTES = G_TILE_M * G_KCHUNK;           // 16 KB
constexpr uint32_t G_STAGE_BYTES = 2 * G_TILE_BYTES;             // A + B
constexpr uint32_t G_TMEM_COLS = 512;

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
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > 4000000000ll) return false;
    }
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
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
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

struct GemmParams {
    uint64_t n_rows;
    uint32_t n_row_tiles, n_q_tiles, q_pad, k_chunks;
    uint32_t *checksum;   // [q_pad]
    int *error_flag;
};

// In-place limb sort of one 128-byte row of the A tile (32 uint32 elements): output 32-byte slice a = byte a of the
// 32 elements.  The row is stored as eight 16-byte chunks, chunk c at physical position c ^ (row & 7) (128-byte swizzle).
__device__ __forceinline__ void limb_sort_row(uint8_t *tile, uint32_t row) {
    uint4 *rowp = reinterpret_cast<uint4 *>(tile + row * 128);
    const uint32_t x = row & 7;
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 v = rowp[c ^ x];
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {   // output chunk 2a + h holds limb a of elements 16h .. 16h+15
            uint32_t o[4];
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int e = 16 * h + 4 * m;
                const uint32_t lo = __byte_perm(w[e], w[e + 1], 0x0040 + a * 0x0011);        // (w[e].b_a, w[e+1].b_a, -, -)
                const uint32_t hi = __byte_perm(w[e + 2], w[e + 3], 0x0040 + a * 0x0011);
                o[m] = __byte_perm(lo, hi, 0x5410);
            }
            rowp[(2 * a + h) ^ x] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

__global__ void __launch_bounds__(G_THREADS, 1) ipgemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b, const GemmParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[G_STAGES], bar_empty[G_STAGES], bar_accum;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < G_STAGES; i++) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_accum, 1);
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
    const uint32_t idesc = umma_idesc_u8(G_TILE_M, G_TILE_N);

    // this CTA's work, flattened: iteration j = ((qt * my_tiles) + r) * k_chunks + kc,  row tile = blockIdx.x + r * gridDim.x
    const uint32_t my_tiles = P.n_row_tiles > blockIdx.x ? (P.n_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t per_qt = my_tiles * P.k_chunks, total = per_qt * P.n_q_tiles;
    auto issue_tma = [&](uint32_t j) -> bool {   // thread 0 only
        const uint32_t stage = j % G_STAGES, ring = (j / G_STAGES) & 1;
        if (j >= (uint32_t)G_STAGES && !mbar_wait(&bar_empty[stage], ring ^ 1)) return false;
        const uint32_t qt = j / per_qt, rem = j % per_qt, rt = blockIdx.x + (rem / P.k_chunks) * gridDim.x, kc = rem % P.k_chunks;
        uint8_t *sa = smem + stage * G_STAGE_BYTES;
        mbar_expect_tx(&bar_full[stage], G_STAGE_BYTES);
        tma_load_2d(sa, &map_a, &bar_full[stage], (int32_t)(kc * G_KCHUNK), (int32_t)(rt * G_TILE_M));
        tma_load_2d(sa + G_TILE_BYTES, &map_b, &bar_full[stage], (int32_t)(kc * G_KCHUNK), (int32_t)(qt * G_TILE_N));
        return true;
    };
    if (threadIdx.x == 0)
        for (uint32_t j = 0; j < (uint32_t)G_PREFETCH && j < total; j++)
            if (!issue_tma(j)) s_abort = 1;
    __syncthreads();

    // The accumulators are NOT drained per row tile: D_s keeps accumulating over all row tiles of a query tile.  s32
    // accumulation wraps mod 2^32 and only sum_i D_s[i][t] << 8s mod 2^32 is wanted, so the wrap is harmless, and the
    // TMEM read-out (256 KB) happens once per query tile instead of once per (query tile, row tile).
    uint32_t accum_phase = 0;
    for (uint32_t j = 0; j < total; j++) {
        const uint32_t stage = j % G_STAGES, ring = (j / G_STAGES) & 1, rem = j % per_qt;
        uint8_t *sa = smem + stage * G_STAGE_BYTES;
        if (threadIdx.x == 0 && j + G_PREFETCH < total && !issue_tma(j + G_PREFETCH)) s_abort = 1;
        if (!mbar_wait(&bar_full[stage], ring)) s_abort = 2;
        limb_sort_row(sa, threadIdx.x);                                     // generic-proxy writes to the A tile ...
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // ... made visible to the tensor core's async proxy
        __syncthreads();
        if (s_abort) break;   // uniform: every writer of s_abort wrote before the barrier
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = smem_u32(sa), b_addr = a_addr + G_TILE_BYTES;
#pragma unroll
            for (uint32_t a = 0; a < 4; a++)
#pragma unroll
                for (uint32_t b = 0; a + b < 4; b++)   // limb pair (a, b) accumulates into D_{a+b}; the first pair of a shift is a = 0
                    umma_i8(tmem_base + (a + b) * G_TILE_N, umma_desc_sw128(a_addr + a * G_UMMA_K), umma_desc_sw128(b_addr + b * G_UMMA_K),
                            idesc, (rem != 0 || a != 0) ? 1u : 0u);
            umma_commit(&bar_empty[stage]);                 // the stage may be refilled once these MMAs have read it
            if (rem + 1 == per_qt) umma_commit(&bar_accum); // every MMA of this query tile has landed in TMEM
        }
        if (rem + 1 != per_qt) continue;
        // ---- end of a query tile: fold the four shifted accumulators, sum over the rows, publish ----
        if (!mbar_wait(&bar_accum, accum_phase)) s_abort = 3;
        accum_phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!s_abort) {  // thread = TMEM lane; warp w owns lanes 32w..32w+31
            const uint32_t lane_addr = tmem_base + ((warp * 32u) << 16), qt = j / per_qt;
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
                    if (lane == 0 && v) atomicAdd(P.checksum + qt * G_TILE_N + c0 + e, v);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM is overwritten by the next query tile's first MMAs
        if (s_abort) break;
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
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint64_t row_bytes) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error(PM_ERR_CUDA, "ip gemm: cuTensorMapEncodeTiled is not available");
    cuuint64_t dims[2] = {row_bytes, rows}, strides[1] = {row_bytes};
    cuuint32_t box[2] = {G_KCHUNK, G_TILE_M}, estr[2] = {1, 1};
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
    if ((rc = make_map(&map_a, db->d_rows, db->n_rows, kbytes))) return rc;
    if ((rc = make_map(&map_b, bmat, q_pad, kbytes))) return rc;
    GemmParams P;
    P.n_rows = db->n_rows;
    P.n_row_tiles = (uint32_t)((db->n_rows + G_TILE_M - 1) / G_TILE_M);
    P.n_q_tiles = q_pad / G_TILE_N;
    P.q_pad = q_pad;
    P.k_chunks = (uint32_t)(kbytes / G_KCHUNK);
    P.checksum = cs_pad;
    P.error_flag = err;
    const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + 1024;
    PM_CUDA(cudaFuncSetAttribute(ipgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<uint32_t>(P.n_row_tiles, (uint32_t)db->sm_count);
    ipgemm_kernel<<<grid, G_THREADS, smem, st>>>(map_a, map_b, P);
    PM_CHECK_LAUNCH();
    count_launch();
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
