// L2Dist in the reference's evaluation order, shared by the distance kernels (pm_ann.cu) and the client's finish kernel
// (pm_client.cuh), which computes the distances of the entries it has just finished.
#pragma once
#include <stdint.h>

namespace pm {

// ---------------------------------------------------------------------------------------------
// A9: L2Dist in the reference's evaluation order (graphann/l2_distance_amd64.s:4-36 and
// build_graph.go:119-127):  8 strided partial sums  acc_l += fl(fl(a-b)^2)  (sub, mul, add rounded
// separately, no FMA), then ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7)), then the dim%8 scalar tail.
// Two lanes evaluate one distance: the even lane owns SIMD lanes 0-3, the odd lane 4-7, each
// streaming its half of every 32-byte step as one 16-byte load; the halves meet in one shuffle.
// ---------------------------------------------------------------------------------------------
template <bool ALIGNED16>
__device__ __forceinline__ float l2_half_pair(const float *__restrict__ a, const float *__restrict__ b, uint32_t dim,
                                              int half) {
    const uint32_t body = dim & ~7u;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *pa = a + 4 * half, *pb = b + 4 * half;
#pragma unroll 4
    for (uint32_t i = 0; i < body; i += 8) {
        float4 x, y;
        if (ALIGNED16) {
            x = *reinterpret_cast<const float4 *>(pa + i);
            y = *reinterpret_cast<const float4 *>(pb + i);
        } else {
            x = make_float4(pa[i], pa[i + 1], pa[i + 2], pa[i + 3]);
            y = make_float4(pb[i], pb[i + 1], pb[i + 2], pb[i + 3]);
        }
        float d0 = __fsub_rn(x.x, y.x), d1 = __fsub_rn(x.y, y.y), d2 = __fsub_rn(x.z, y.z), d3 = __fsub_rn(x.w, y.w);
        a0 = __fadd_rn(a0, __fmul_rn(d0, d0));
        a1 = __fadd_rn(a1, __fmul_rn(d1, d1));
        a2 = __fadd_rn(a2, __fmul_rn(d2, d2));
        a3 = __fadd_rn(a3, __fmul_rn(d3, d3));
    }
    return __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
}
// full distance for the lane pair (both lanes return the same value)
template <bool ALIGNED16>
__device__ __forceinline__ float l2_pair(const float *a, const float *b, uint32_t dim, int half) {
    float h = l2_half_pair<ALIGNED16>(a, b, dim, half);
    float o = __shfl_xor_sync(0xffffffffu, h, 1);
    float d = half ? __fadd_rn(o, h) : __fadd_rn(h, o);  // lo + hi on both lanes
    for (uint32_t i = dim & ~7u; i < dim; i++) {
        float x = __fsub_rn(a[i], b[i]);
        d = __fadd_rn(d, __fmul_rn(x, x));
    }
    return d;
}

}  // namespace pm
