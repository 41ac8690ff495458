// Internal helpers shared by the kernels of libpacmann_cuda.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/pacmann_cuda.h"

#define PM_EXPORT extern "C" __attribute__((visibility("default")))

struct pm_db {
    int device;
    uint64_t n_rows, entry_u64;
    uint64_t *d_rows;          // [n_rows][entry_u64], 256-byte aligned (cudaMalloc)
    cudaStream_t stream;       // compute stream of this handle
    cudaStream_t copy_stream;  // D2H / H2D overlap stream
    cudaEvent_t ev[4];
    // grow-only device scratch (per handle, so handles stay re-entrant with respect to each other)
    void *scratch[4];
    size_t scratch_bytes[4];
    std::mutex mu;             // serialises calls on one handle
    int sm_count;
    bool owns_rows;            // false for pm_db_wrap handles
    unsigned int *sync_pool;   // 64 barrier counters (one 128-byte line each), handed out round-robin
    std::atomic<unsigned> sync_next;
};

namespace pm {

int set_error(int code, const char *fmt, ...);
// launch-time tuning knobs (pm_tuning_set / environment), see pm_core.cu
enum Tune { T_HG_SYNC, T_HG_WARPS, T_HG_NTAB, T_HG_TAIL_SPLIT, T_HG_SERPENTINE, T_HG_XBYTES, T_ANS_SPLIT, T_HG_D2H_GROUPS, T_SEARCH_ANS_STREAM, T_COUNT };
int tune(Tune t);
void count_launch(uint64_t n = 1);
int ensure_device(int device);  // cudaSetDevice + one-time table upload; returns PM_OK / error
int sm_count(int device);
const void *zero_page(int device);  // 4 KB of device zeros, allocated once per device
// tensor-core path of the uint32 inner-product scan (pm_ipgemm.cu)
bool ipgemm_applicable(const pm_db *db, uint64_t dim, uint64_t nq, const uint32_t *ip_out);
size_t ipgemm_scratch_bytes(uint64_t dim, uint64_t nq);
int ipgemm_enqueue(pm_db *db, uint64_t dim, const uint32_t *queries, uint64_t nq, uint32_t *checksum, void *scratch_dev, cudaStream_t st);
int ipgemm_check(void *scratch_dev, uint64_t dim, uint64_t nq);
// grow-only scratch slot on a handle
int scratch(pm_db *db, int slot, size_t bytes, void **out);
unsigned int *sync_counter(pm_db *db);  // next barrier counter of the handle's pool
// per-device grow-only workspace + stream for the handle-less entry points (pm_l2_*, pm_prf_batch, ...); the returned
// lock serialises those calls on one device and must be held until the stream has been synchronised
int dev_work(int device, size_t bytes, void **ptr, cudaStream_t *stream, std::unique_lock<std::mutex> *lock);

#define PM_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return pm::set_error(PM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                                 __FILE__, __LINE__);                                                   \
    } while (0)

#define PM_CHECK_LAUNCH()                                                                               \
    do {                                                                                                \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return pm::set_error(PM_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                                   \
    } while (0)

// 128-bit streaming load that does not allocate in L1 (rows are touched once per SM)
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2 *p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// predicated form: zeros when !pred, and no access at all
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

__device__ __forceinline__ void vxor(uint4 &a, const uint4 &b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; }
__device__ __forceinline__ void vxor(uint2 &a, const uint2 &b) { a.x ^= b.x; a.y ^= b.y; }
__device__ __forceinline__ void vzero(uint4 &a) { a = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void vzero(uint2 &a) { a = make_uint2(0, 0); }

}  // namespace pm
