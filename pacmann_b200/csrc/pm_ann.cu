// graphann distance kernels for sm_100a: bit-exact fp32 L2 (A9) and the uint32 inner-product linear
// scan (A11).  See include/pacmann_cuda.h for the reference seams.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "pm_common.cuh"
#include "pm_l2.cuh"

namespace pm {

constexpr int L2_THREADS = 128;
// grid.x = query, grid.y = candidate block; query vector staged in shared memory
template <bool ALIGNED16>
__global__ void __launch_bounds__(L2_THREADS) l2_batch_kernel(const uint64_t *db, uint64_t n_rows, uint64_t entry_u64,
                                                              uint32_t dim, const float *queries, const int64_t *ids,
                                                              uint64_t k, float *out) {
    extern __shared__ __align__(16) float s_q[];
    const uint64_t q = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < dim; i += L2_THREADS) s_q[i] = queries[q * dim + i];
    __syncthreads();
    const int half = threadIdx.x & 1;
    const uint64_t per_block = L2_THREADS / 2;
    for (uint64_t j0 = (uint64_t)blockIdx.y * per_block; j0 < k; j0 += (uint64_t)gridDim.y * per_block) {
        const uint64_t j = j0 + (threadIdx.x >> 1);
        const bool in = j < k;
        const int64_t id = in ? ids[q * k + j] : -1;
        const bool ok = in && id >= 0 && (uint64_t)id < n_rows;
        const float *row = reinterpret_cast<const float *>(db + (ok ? (uint64_t)id : 0) * entry_u64);
        float d = l2_pair<ALIGNED16>(row, s_q, dim, half);
        if (in && half == 0) out[q * k + j] = ok ? d : INFINITY;
    }
}
// b row of pair i: b + i*b_stride, or b + b_index[i]*b_stride when b_index is given (several queries in one call)
__global__ void __launch_bounds__(L2_THREADS) l2_pairs_kernel(const float *a, uint64_t a_stride, const float *b, uint64_t b_stride,
                                                              const uint32_t *b_index, uint64_t n, uint32_t dim, float *out) {
    const int half = threadIdx.x & 1;
    const uint64_t stride = (uint64_t)gridDim.x * (L2_THREADS / 2);
    const uint64_t n_up = (n + 15) & ~15ull;  // keep whole warps in the shuffle
    for (uint64_t i = (uint64_t)blockIdx.x * (L2_THREADS / 2) + (threadIdx.x >> 1); i < n_up; i += stride) {
        const bool ok = i < n;
        const uint64_t r = ok ? i : 0;
        float d = l2_pair<false>(a + r * a_stride, b + (b_index ? (uint64_t)b_index[r] : r) * b_stride, dim, half);
        if (ok && half == 0) out[i] = d;
    }
}

// out[p] = L2Dist(vector of row ids_a[p], vector of row ids_b[p]) over the resident table: the distance matrices of
// graph construction (robustPrune, build_graph.go:169-236: u -> candidates and candidate x candidate)
template <bool ALIGNED16>
__global__ void __launch_bounds__(L2_THREADS) l2_idpairs_kernel(const uint64_t *db, uint64_t n_rows, uint64_t entry_u64, uint32_t dim,
                                                                const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out) {
    const int half = threadIdx.x & 1;
    const uint64_t stride = (uint64_t)gridDim.x * (L2_THREADS / 2);
    const uint64_t n_up = (n + 15) & ~15ull;  // keep whole warps in the shuffle
    for (uint64_t i = (uint64_t)blockIdx.x * (L2_THREADS / 2) + (threadIdx.x >> 1); i < n_up; i += stride) {
        const bool in = i < n;
        const int64_t a = in ? ids_a[i] : -1, b = in ? ids_b[i] : -1;
        const bool ok = in && a >= 0 && b >= 0 && (uint64_t)a < n_rows && (uint64_t)b < n_rows;
        const float *ra = reinterpret_cast<const float *>(db + (ok ? (uint64_t)a : 0) * entry_u64);
        const float *rb = reinterpret_cast<const float *>(db + (ok ? (uint64_t)b : 0) * entry_u64);
        float d = l2_pair<ALIGNED16>(ra, rb, dim, half);
        if (in && half == 0) out[i] = ok ? d : INFINITY;
    }
}

// ---------------------------------------------------------------------------------------------
// A11: uint32 wrapping inner-product scan (graphann/l2_distance_amd64.s:39-68, graphann_test.go:268-273)
// Tiles of IP_ROWS consecutive rows are contiguous in memory: they are staged into shared memory with fully
// coalesced 16-byte cp.async copies (row stride padded by one vector so the per-row reads below are bank-conflict
// free), then one thread per row accumulates QT queries at a time; query words arrive as warp-wide broadcasts.
// Two CTAs per SM overlap one CTA's copy phase with the other's multiply phase.  Integer arithmetic mod 2^32 is
// associative, so any summation order gives the reference's result bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int IP_THREADS = 128;
constexpr int IP_ROWS = 128;
constexpr int IP_QT_MAX = 16;
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
template <int QT>
__global__ void __launch_bounds__(IP_THREADS, 2) ip_scan_kernel(const uint4 *rows, uint64_t n_rows, uint32_t dim4, uint32_t tile_rows,
                                                                const uint32_t *queries, uint32_t n_queries, uint32_t q0,
                                                                uint32_t *checksum, uint32_t *ip_out) {
    extern __shared__ __align__(16) uint32_t s_qw[];
    uint4 *s_q4 = reinterpret_cast<uint4 *>(s_qw);           // [QT][dim4]
    uint4 *s_tile = s_q4 + QT * dim4;                        // [tile_rows][dim4 + 1]
    const uint32_t nq = min((uint32_t)QT, n_queries - q0), pitch = dim4 + 1;
    for (uint32_t i = threadIdx.x; i < QT * dim4; i += IP_THREADS) {
        uint32_t t = i / dim4, c = i % dim4;
        s_q4[i] = t < nq ? reinterpret_cast<const uint4 *>(queries)[(uint64_t)(q0 + t) * dim4 + c] : make_uint4(0, 0, 0, 0);
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t total[QT];
#pragma unroll
    for (int t = 0; t < QT; t++) total[t] = 0;
    const uint64_t n_tiles = (n_rows + tile_rows - 1) / tile_rows;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t r0 = tile * tile_rows;
        const uint32_t nr = (uint32_t)min((uint64_t)tile_rows, n_rows - r0);
        __syncthreads();  // previous tile fully consumed (and the queries written, first time round)
        for (uint32_t r = warp; r < nr; r += IP_THREADS / 32)
            for (uint32_t c = lane; c < dim4; c += 32) cp_async16(s_tile + r * pitch + c, rows + (r0 + r) * dim4 + c);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (threadIdx.x < nr) {
            uint32_t acc[QT];
#pragma unroll
            for (int t = 0; t < QT; t++) acc[t] = 0;
            const uint4 *row = s_tile + threadIdx.x * pitch;
#pragma unroll 4
            for (uint32_t c = 0; c < dim4; c++) {
                const uint4 v = row[c];
#pragma unroll
                for (int t = 0; t < QT; t++) {
                    const uint4 qv = s_q4[t * dim4 + c];
                    acc[t] += v.x * qv.x + v.y * qv.y + v.z * qv.z + v.w * qv.w;
                }
            }
#pragma unroll
            for (int t = 0; t < QT; t++) {
                total[t] += acc[t];
                if (ip_out && (uint32_t)t < nq) ip_out[(uint64_t)(q0 + t) * n_rows + r0 + threadIdx.x] = acc[t];
            }
        }
    }
    // block reduction, then one atomic per query per block
    __shared__ uint32_t s_red[IP_THREADS / 32][IP_QT_MAX];
#pragma unroll
    for (int t = 0; t < QT; t++) {
        uint32_t v = total[t];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < nq) {
        uint32_t v = 0;
        for (int w = 0; w < IP_THREADS / 32; w++) v += s_red[w][threadIdx.x];
        atomicAdd(checksum + q0 + threadIdx.x, v);
    }
}

// Few queries, checksums only (the reference's scan loop keeps nothing but the sum, graphann_test.go:268-273): a pure
// stream.  Thread t walks the table as 16-byte vectors t, t + T, t + 2T, ... (T = all threads: fully coalesced, no shared
// memory staging, eight loads in flight per thread); its column index advances by T mod dim4 with one conditional
// subtract.  Integer sums mod 2^32 are order-free, so the result is the reference's bit for bit.
constexpr int IPS_THREADS = 256, IPS_UNROLL = 8;
template <int QT>
__global__ void __launch_bounds__(IPS_THREADS) ip_stream_kernel(const uint4 *__restrict__ rows, uint64_t n_vec, uint32_t dim4,
                                                                const uint32_t *queries, uint32_t n_queries, uint32_t q0, uint32_t *checksum) {
    extern __shared__ __align__(16) uint32_t s_qw[];
    uint4 *s_q4 = reinterpret_cast<uint4 *>(s_qw);           // [QT][dim4]
    const uint32_t nq = min((uint32_t)QT, n_queries - q0);
    for (uint32_t i = threadIdx.x; i < QT * dim4; i += IPS_THREADS) {
        const uint32_t t = i / dim4, c = i % dim4;
        s_q4[i] = t < nq ? reinterpret_cast<const uint4 *>(queries)[(uint64_t)(q0 + t) * dim4 + c] : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint64_t T = (uint64_t)gridDim.x * IPS_THREADS;
    const uint32_t step = (uint32_t)(T % dim4);
    uint64_t idx = (uint64_t)blockIdx.x * IPS_THREADS + threadIdx.x;
    uint32_t c = (uint32_t)(idx % dim4);
    uint32_t total[QT];
#pragma unroll
    for (int t = 0; t < QT; t++) total[t] = 0;
    while (idx < n_vec) {
        uint4 v[IPS_UNROLL];
        uint32_t cc[IPS_UNROLL];
#pragma unroll
        for (int u = 0; u < IPS_UNROLL; u++) {
            const uint64_t i = idx + (uint64_t)u * T;
            cc[u] = c;
            v[u] = i < n_vec ? ldg_stream(rows + i) : make_uint4(0, 0, 0, 0);
            c += step;
            if (c >= dim4) c -= dim4;
        }
#pragma unroll
        for (int u = 0; u < IPS_UNROLL; u++)
#pragma unroll
            for (int t = 0; t < QT; t++) {
                const uint4 qv = s_q4[t * dim4 + cc[u]];
                total[t] += v[u].x * qv.x + v[u].y * qv.y + v[u].z * qv.z + v[u].w * qv.w;
            }
        idx += (uint64_t)IPS_UNROLL * T;
    }
    __shared__ uint32_t s_red[IPS_THREADS / 32][QT];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < QT; t++) {
        uint32_t v = total[t];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < nq) {
        uint32_t v = 0;
        for (int w = 0; w < IPS_THREADS / 32; w++) v += s_red[w][threadIdx.x];
        atomicAdd(checksum + q0 + threadIdx.x, v);
    }
}

// out[i] = L2Dist(a + i*a_stride, b + i*b_stride) for device-resident rows (strides in floats; b_stride 0 = one query)
int l2_batch_enqueue(pm_db *db, uint64_t dim, const float *queries, uint64_t nq, const int64_t *ids, uint64_t k, float *out,
                     cudaStream_t st) {
    if (nq == 0 || k == 0) return PM_OK;
    if (dim * 4 > db->entry_u64 * 8) return set_error(PM_ERR_ARG, "l2: dim %llu does not fit in a row", (unsigned long long)dim);
    if (dim > 8192) return set_error(PM_ERR_UNSUPPORTED, "l2: dim too large");
    if (nq > 0x7fffffffull) return set_error(PM_ERR_UNSUPPORTED, "l2: too many queries in one call");
    uint64_t by = (k + L2_THREADS / 2 - 1) / (L2_THREADS / 2);
    if (by > 65535) by = 65535;
    dim3 grid((unsigned)nq, (unsigned)by);
    const size_t smem = ((dim + 3) & ~3ull) * 4;
    if (db->entry_u64 % 2 == 0)
        l2_batch_kernel<true><<<grid, L2_THREADS, smem, st>>>(db->d_rows, db->n_rows, db->entry_u64, (uint32_t)dim, queries, ids, k, out);
    else
        l2_batch_kernel<false><<<grid, L2_THREADS, smem, st>>>(db->d_rows, db->n_rows, db->entry_u64, (uint32_t)dim, queries, ids, k, out);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}

int ip_scan_enqueue(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t nq, uint32_t *checksum, uint32_t *ip_out,
                    cudaStream_t st) {
    if (nq == 0) return PM_OK;
    if (dim == 0 || dim % 16) return set_error(PM_ERR_ARG, "ip scan: dim must be a positive multiple of 16 (reference loop)");
    if (dim != db->entry_u64 * 2) return set_error(PM_ERR_ARG, "ip scan: dim must equal 2*entry_u64");
    if (dim > 2048) return set_error(PM_ERR_UNSUPPORTED, "ip scan: dim too large");
    if (nq > 0x7fffffffull) return set_error(PM_ERR_UNSUPPORTED, "ip scan: too many queries");
    if (ipgemm_applicable(db, dim, nq, ip_out)) {  // many queries, checksums only: int8 limb GEMM on the tensor cores
        void *ws = nullptr;
        int rc = scratch(db, 2, ipgemm_scratch_bytes(dim, nq), &ws);
        if (rc) return rc;
        return ipgemm_enqueue(db, dim, queries, nq, checksum, ws, st);
    }
    PM_CUDA(cudaMemsetAsync(checksum, 0, nq * 4, st));
    const uint32_t dim4 = (uint32_t)(dim / 4);
    // rows per tile: as many as fit beside the queries in ~100 KB, so that two CTAs share an SM
    const size_t q_bytes = (size_t)IP_QT_MAX * dim * 4, row_bytes = (size_t)(dim4 + 1) * 16;
    uint32_t tile_rows = (uint32_t)std::min<size_t>(IP_ROWS, (100 * 1024 - std::min<size_t>(q_bytes, 48 * 1024)) / row_bytes);
    if (tile_rows == 0) return set_error(PM_ERR_UNSUPPORTED, "ip scan: dim too large");
    const uint64_t n_tiles = (db->n_rows + tile_rows - 1) / tile_rows;
    uint64_t blocks = std::min<uint64_t>(n_tiles ? n_tiles : 1, (uint64_t)db->sm_count * 2);
    auto launch = [&](auto kern, int qt, uint64_t q0) -> int {
        const size_t smem = (size_t)qt * dim * 4 + (size_t)tile_rows * row_bytes;
        PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, IP_THREADS, smem, st>>>((const uint4 *)db->d_rows, db->n_rows, dim4, tile_rows, queries, (uint32_t)nq,
                                                        (uint32_t)q0, checksum, ip_out);
        PM_CHECK_LAUNCH();
        count_launch();
        return PM_OK;
    };
    uint64_t q0 = 0;
    int rc = PM_OK;
    if (!ip_out && nq <= 4) {   // checksums of a few queries: the streaming kernel (no staging), one pass
        const uint64_t n_vec = db->n_rows * dim4;
        const unsigned sblocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (n_vec + IPS_THREADS * IPS_UNROLL - 1) / (IPS_THREADS * IPS_UNROLL)),
                                                               (uint64_t)db->sm_count * 8);
        const size_t smem = (size_t)(nq == 1 ? 1 : 4) * dim * 4;
        if (nq == 1) ip_stream_kernel<1><<<sblocks, IPS_THREADS, smem, st>>>((const uint4 *)db->d_rows, n_vec, dim4, queries, (uint32_t)nq, 0, checksum);
        else ip_stream_kernel<4><<<sblocks, IPS_THREADS, smem, st>>>((const uint4 *)db->d_rows, n_vec, dim4, queries, (uint32_t)nq, 0, checksum);
        PM_CHECK_LAUNCH();
        count_launch();
        return PM_OK;
    }
    while (q0 < nq && rc == PM_OK) {  // passes of 16 queries, then 4, then single queries for the remainder
        const uint64_t left = nq - q0;
        if (left >= 16 || left > 4) { rc = launch(ip_scan_kernel<16>, 16, q0); q0 += 16; }
        else if (left > 1) { rc = launch(ip_scan_kernel<4>, 4, q0); q0 += 4; }
        else { rc = launch(ip_scan_kernel<1>, 1, q0); q0 += 1; }
    }
    return rc;
}

}  // namespace pm

using namespace pm;

// b_rows == n: pairwise; b_rows == 1: one query against n vectors
static int l2_host(const float *a, const float *b, uint64_t b_rows, uint64_t n, uint64_t dim, float *out, int device) {
    if (n && (!a || !b || !out)) return set_error(PM_ERR_ARG, "pm_l2: null pointer");
    if (n == 0) return PM_OK;
    if (dim > 0xffffffffull) return set_error(PM_ERR_UNSUPPORTED, "pm_l2: dim too large");
    int rc = ensure_device(device);
    if (rc) return rc;
    const size_t ab = n * dim * 4, bb = b_rows * dim * 4;
    void *w = nullptr;
    cudaStream_t st;
    std::unique_lock<std::mutex> lock;
    if ((rc = dev_work(device, ab + bb + n * 4 + 256, &w, &st, &lock))) return rc;
    float *d = (float *)w, *d_b = d + n * dim, *d_out = d_b + b_rows * dim;
    PM_CUDA(cudaMemcpyAsync(d, a, ab, cudaMemcpyHostToDevice, st));
    PM_CUDA(cudaMemcpyAsync(d_b, b, bb, cudaMemcpyHostToDevice, st));
    uint64_t blocks = (n + L2_THREADS / 2 - 1) / (L2_THREADS / 2);
    if (blocks > 148 * 16) blocks = 148 * 16;
    l2_pairs_kernel<<<(unsigned)blocks, L2_THREADS, 0, st>>>(d, dim, d_b, b_rows == 1 ? 0 : dim, nullptr, n, (uint32_t)dim, d_out);
    PM_CHECK_LAUNCH();
    count_launch();
    PM_CUDA(cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaStreamSynchronize(st));
    return PM_OK;
}
PM_EXPORT int pm_l2_pairs(const float *a, const float *b, uint64_t n, uint64_t dim, float *out, int device) {
    return l2_host(a, b, n, n, dim, out, device);
}
PM_EXPORT int pm_l2_query(const float *vecs, uint64_t n, uint64_t dim, const float *query, float *out, int device) {
    return l2_host(vecs, query, 1, n, dim, out, device);
}

PM_EXPORT int pm_l2_batch_dev(pm_db *db, uint64_t dim, const float *queries, uint64_t n_queries, const int64_t *ids, uint64_t k,
                              float *out, void *stream) {
    if (!db || (n_queries && k && (!queries || !ids || !out))) return set_error(PM_ERR_ARG, "pm_l2_batch_dev: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    return l2_batch_enqueue(db, dim, queries, n_queries, ids, k, out, stream ? (cudaStream_t)stream : db->stream);
}

PM_EXPORT int pm_l2_batch(pm_db *db, uint64_t dim, const float *queries, uint64_t n_queries, const int64_t *ids, uint64_t k,
                          float *out) {
    if (!db || (n_queries && k && (!queries || !ids || !out))) return set_error(PM_ERR_ARG, "pm_l2_batch: null pointer");
    if (n_queries == 0 || k == 0) return PM_OK;
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const size_t qb = n_queries * dim * 4, ib = n_queries * k * 8, ob = n_queries * k * 4;
    void *d_in = nullptr, *d_out = nullptr;
    if ((rc = scratch(db, 1, ib + qb, &d_in))) return rc;
    if ((rc = scratch(db, 0, ob, &d_out))) return rc;
    int64_t *d_ids = (int64_t *)d_in;
    float *d_q = (float *)((char *)d_in + ib);
    PM_CUDA(cudaMemcpyAsync(d_ids, ids, ib, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_q, queries, qb, cudaMemcpyHostToDevice, db->stream));
    if ((rc = l2_batch_enqueue(db, dim, d_q, n_queries, d_ids, k, (float *)d_out, db->stream))) return rc;
    PM_CUDA(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

static int l2_idpairs_enqueue(pm_db *db, uint64_t dim, const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out, cudaStream_t st) {
    if (n == 0) return PM_OK;
    if (dim * 4 > db->entry_u64 * 8) return set_error(PM_ERR_ARG, "l2: dim %llu does not fit in a row", (unsigned long long)dim);
    if (dim > 8192) return set_error(PM_ERR_UNSUPPORTED, "l2: dim too large");
    uint64_t blocks = (n + L2_THREADS / 2 - 1) / (L2_THREADS / 2);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (db->entry_u64 % 2 == 0)
        l2_idpairs_kernel<true><<<(unsigned)blocks, L2_THREADS, 0, st>>>(db->d_rows, db->n_rows, db->entry_u64, (uint32_t)dim, ids_a, ids_b, n, out);
    else
        l2_idpairs_kernel<false><<<(unsigned)blocks, L2_THREADS, 0, st>>>(db->d_rows, db->n_rows, db->entry_u64, (uint32_t)dim, ids_a, ids_b, n, out);
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
PM_EXPORT int pm_l2_idpairs_dev(pm_db *db, uint64_t dim, const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out, void *stream) {
    if (!db || (n && (!ids_a || !ids_b || !out))) return set_error(PM_ERR_ARG, "pm_l2_idpairs_dev: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    return l2_idpairs_enqueue(db, dim, ids_a, ids_b, n, out, stream ? (cudaStream_t)stream : db->stream);
}
PM_EXPORT int pm_l2_idpairs(pm_db *db, uint64_t dim, const int64_t *ids_a, const int64_t *ids_b, uint64_t n, float *out) {
    if (!db || (n && (!ids_a || !ids_b || !out))) return set_error(PM_ERR_ARG, "pm_l2_idpairs: null pointer");
    if (n == 0) return PM_OK;
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    void *d_in = nullptr, *d_out = nullptr;
    if ((rc = scratch(db, 1, 2 * n * 8, &d_in))) return rc;
    if ((rc = scratch(db, 0, n * 4, &d_out))) return rc;
    int64_t *d_a = (int64_t *)d_in, *d_b = d_a + n;
    PM_CUDA(cudaMemcpyAsync(d_a, ids_a, n * 8, cudaMemcpyHostToDevice, db->stream));
    PM_CUDA(cudaMemcpyAsync(d_b, ids_b, n * 8, cudaMemcpyHostToDevice, db->stream));
    if ((rc = l2_idpairs_enqueue(db, dim, d_a, d_b, n, (float *)d_out, db->stream))) return rc;
    PM_CUDA(cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    return PM_OK;
}

PM_EXPORT int pm_ip_u32_scan_dev(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t n_queries, uint32_t *checksum_out,
                                 uint32_t *ip_out, void *stream) {
    if (!db || (n_queries && (!queries || !checksum_out))) return set_error(PM_ERR_ARG, "pm_ip_u32_scan_dev: null pointer");
    int rc = ensure_device(db->device);
    if (rc) return rc;
    return ip_scan_enqueue(db, dim, queries, n_queries, checksum_out, ip_out, stream ? (cudaStream_t)stream : db->stream);
}

PM_EXPORT int pm_ip_u32_scan(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t n_queries, uint32_t *checksum_out,
                             uint32_t *ip_out) {
    if (!db || (n_queries && (!queries || !checksum_out))) return set_error(PM_ERR_ARG, "pm_ip_u32_scan: null pointer");
    if (n_queries == 0) return PM_OK;
    int rc = ensure_device(db->device);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const size_t qb = n_queries * dim * 4, cb = n_queries * 4, pb = ip_out ? n_queries * db->n_rows * 4 : 0;
    void *d_in = nullptr, *d_out = nullptr;
    if ((rc = scratch(db, 1, qb, &d_in))) return rc;
    if ((rc = scratch(db, 0, ((cb + 255) & ~255ull) + pb, &d_out))) return rc;
    uint32_t *d_ip = pb ? (uint32_t *)((char *)d_out + ((cb + 255) & ~255ull)) : nullptr;
    PM_CUDA(cudaMemcpyAsync(d_in, queries, qb, cudaMemcpyHostToDevice, db->stream));
    if ((rc = ip_scan_enqueue(db, dim, (const uint32_t *)d_in, n_queries, (uint32_t *)d_out, d_ip, db->stream))) return rc;
    PM_CUDA(cudaMemcpyAsync(checksum_out, d_out, cb, cudaMemcpyDeviceToHost, db->stream));
    if (pb) PM_CUDA(cudaMemcpyAsync(ip_out, d_ip, pb, cudaMemcpyDeviceToHost, db->stream));
    PM_CUDA(cudaStreamSynchronize(db->stream));
    if (ipgemm_applicable(db, dim, n_queries, ip_out)) return ipgemm_check(db->scratch[2], dim, n_queries);
    return PM_OK;
}
