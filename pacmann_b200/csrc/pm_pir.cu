// PianoPIR kernels for sm_100a: AES-PRF, key schedule, hint generation (A5/A7), server answer
// (A6/A8), row gather, xorSlices.  See include/pacmann_cuda.h for the reference seams.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pm_aes.cuh"
#include <cooperative_groups.h>

#include "pm_common.cuh"
#include "pm_hg_params.cuh"

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
    int rc;   // the hint kernel's translation units keep their own copy
    if ((rc = hg_upload_tables_a(te0)) || (rc = hg_upload_tables_b(te0)) || (rc = hg_upload_tables_c(te0))) return rc;
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
int hintgen_enqueue(pm_db *db, const pm_hint_job *jobs, uint64_t n_jobs, cudaStream_t st);   // pm_hintgen.cu

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
    else if (x == 40) PM_ANS(8, 5);      // 640-byte rows (SIFT-shaped): 8 lanes x 5 vectors keep every lane busy (32 x 2 idles 37 %)
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
    // 96 sub-queries, us: 1 CTA 24.9 | 2 21.2 | 3 18.4 | 4 16.8 | 6 16.5 | 8 20.0).  The ans_split knob forces.
    const int force_split = tune(T_ANS_SPLIT);
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
#include "pm_search.cuh"
