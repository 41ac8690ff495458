// A5/A7 hint generation kernel (PianoPIRClient.Preprocessing / UpdatePreprocessing, pianopir/pir.go:267-352, and the
// batch fan-out of pianopir/batch-pir.go:119-155).  Template + launch helpers; instantiated by pm_hg_*.cu (one
// translation unit per PRF variant so that the ~200 instantiations compile in parallel).
//
// Hint-stationary: a group of G lanes owns one hint and keeps its parity in registers across all set_size chunks; each
// lane of the group evaluates the PRF for a different chunk (G chunks per step, all 32 lanes of the warp busy in AES),
// offsets are exchanged by shuffle, and the G lanes then read the selected row together (G*16 contiguous bytes per load
// instruction, whole 128-byte lines for G = 8).  Parities are written once.
//
// Schedule: persistent grid, one CTA per SM, tiles of (warps * 32/G) hints.
//   * full rounds: in round r CTA b owns tile r*grid + b and sweeps its sub-PIR's DB slice; with the round barrier
//     (cooperative launch) all CTAs start every sweep together, so a chunk is fetched from HBM once per round and re-hit
//     in L2 by the other CTAs.  Consecutive rounds sweep in opposite directions ("serpentine"): the end of one sweep is
//     still in L2 when the next one starts there.
//   * the last, partial round is shared stream-K style: its T tiles are T sweeps of work, cut into `grid` equal slices;
//     a CTA XORs the rows of its slice (a range of chunk groups of one tile, or the end of one tile and the start of the
//     next) into the pre-zeroed output with red.global.xor -- no round is ever mostly idle (8 GPUs: 5.16 rounds of work
//     used to cost 6).
#pragma once
#include <algorithm>
#include <cstring>

#include "pm_aes.cuh"
#include "pm_common.cuh"
#include "pm_hg_params.cuh"

#ifndef PM_HG_SWEEP_INLINE
#define PM_HG_SWEEP_INLINE __forceinline__
#endif

namespace pm {

struct RkOfJob {
    const HintParams &P;
    int j;
    __device__ __forceinline__ uint32_t operator[](int i) const { return P.jobs[j].rk[i]; }
};

// XOR into global memory (the destination may be a peer GPU's table mapped over NVLink, hence system scope)
__device__ __forceinline__ void red_xor(uint4 *p, const uint4 &v) {
    asm volatile("red.relaxed.sys.global.xor.b64 [%0], %1;" ::"l"(p), "l"((uint64_t)v.x | ((uint64_t)v.y << 32)) : "memory");
    asm volatile("red.relaxed.sys.global.xor.b64 [%0], %1;" ::"l"((char *)p + 8), "l"((uint64_t)v.z | ((uint64_t)v.w << 32)) : "memory");
}
__device__ __forceinline__ void red_xor(uint2 *p, const uint2 &v) {
    asm volatile("red.relaxed.sys.global.xor.b64 [%0], %1;" ::"l"(p), "l"((uint64_t)v.x | ((uint64_t)v.y << 32)) : "memory");
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
template <typename VT, int G, int NV, int NTAB, int NB, int XB, int U, bool FULL, int PH, int NPH, typename RK>
__device__ __forceinline__ void hg_phase(const AesTab<NTAB> &T, const RK &R, const PrfHint<XB> &g, uint32_t c_next,
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
    prf_rounds<XB, NTAB, NB, PrfPhase<PH, NPH, XB>::first, PrfPhase<PH, NPH, XB>::last>(T, R, g, c_next, st);
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

// One sweep of one tile over the chunk groups [g_begin, g_end) of its sub-PIR (all of them in a full round), forwards or
// backwards; `accumulate` = XOR the result into the pre-zeroed output instead of storing it (shared last round).
template <typename VT, int G, int NV, int NTAB, int NB, int XB, int U, bool FULL>
__device__ PM_HG_SWEEP_INLINE void hg_sweep(const HintParams &P, const AesTab<NTAB> &T, uint32_t tile, uint32_t frac_a, uint32_t frac_b,
                                         bool rev) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GPW = 32 / G, NPH = G / U;
    const int gl = lane & (G - 1), gbase = lane & ~(G - 1), gw = lane / G;
    const uint32_t hints_per_tile = (blockDim.x / 32) * GPW;
    const uint32_t ev = P.ev, evx = P.evx;
    int j = 0;
    while (j + 1 < (int)P.n_jobs && tile >= P.jobs[j + 1].tile_begin) j++;
    const HintJobDev &J = P.jobs[j];
    const RkOfJob R{P, j};
    const uint32_t S = J.set_size, cmask = J.chunk_mask, cshift = J.chunk_shift;
    const uint32_t n_groups = (S + G - 1) / G;
    const uint32_t g_begin = (uint32_t)((uint64_t)frac_a * n_groups / HG_TAIL_K), g_end = (uint32_t)((uint64_t)frac_b * n_groups / HG_TAIL_K);
    if (g_begin >= g_end) return;
    const bool accumulate = !(g_begin == 0 && g_end == n_groups);
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
    const PrfHint<XB> g = prf_hint_part<XB>(T, R, tag);
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
    const uint32_t c_step = rev ? (uint32_t)(-G) : (uint32_t)G;
    uint32_t c_cur = (rev ? g_end - 1 : g_begin) * G + gl;   // this lane's chunk of the current group
    // optional side output: the offset of every (hint, chunk) this sweep evaluates, row-major per hint (16 contiguous bytes
    // per lane group and step); the resident client builds its offset index from it
    uint16_t *off_row = (J.off && active) ? J.off + i * ((S + 7) & ~7u) : nullptr;
    PrfState st;
    prf_rounds<XB, NTAB, NB, 1, 10>(T, R, g, c_cur, st);  // prologue: rows of the first group
    uint32_t row_next = to_row(c_cur, st.s0);
    if (off_row && c_cur < S) off_row[c_cur] = (uint16_t)(st.s0 & cmask);
    for (uint32_t it = g_begin; it < g_end; it++) {
        // c_next runs one group past the range in the last iteration: that PRF value is never used
        const uint32_t row_cur = row_next, c_next = c_cur + c_step;
        hg_phase<VT, G, NV, NTAB, NB, XB, U, FULL, 0, NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
        if (NPH > 1) hg_phase<VT, G, NV, NTAB, NB, XB, U, FULL, (NPH > 1 ? 1 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
        if (NPH > 2) {
            hg_phase<VT, G, NV, NTAB, NB, XB, U, FULL, (NPH > 2 ? 2 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
            hg_phase<VT, G, NV, NTAB, NB, XB, U, FULL, (NPH > 2 ? 3 : 0), NPH>(T, R, g, c_next, st, row_cur, gbase, gl, base, ev, evx, par);
        }
        row_next = to_row(c_next, st.s0);
        if (off_row && it + 1 < g_end && c_next < S) off_row[c_next] = (uint16_t)(st.s0 & cmask);
        c_cur = c_next;
    }
    if (!active) return;
    VT *o = reinterpret_cast<VT *>(J.out) + i * ev + gl;
    if (accumulate) {   // words past evx are never xored and the output is already zero there
#pragma unroll
        for (int k = 0; k < NV; k++)
            if ((uint32_t)(k * G + gl) < evx) red_xor(o + k * G, par[k]);
        return;
    }
#pragma unroll
    for (int k = 0; k < NV; k++)
        if ((uint32_t)(k * G + gl) < ev) o[k * G] = par[k];
    VT z;
    vzero(z);
    for (uint32_t col = NV * G + gl; col < ev; col += G) o[col - gl] = z;
}

template <typename VT, int G, int NV, int NTAB, int NB, int XB, int U, bool FULL>
__global__ void __launch_bounds__(HG_MAX_THREADS, 1) hintgen_kernel(const __grid_constant__ HintParams P) {
    extern __shared__ uint32_t smem[];
    aes_tab_fill<NTAB>(smem, PM_HG_TE0);
    __syncthreads();
    const AesTab<NTAB> T{smem + (threadIdx.x & 31)};
    const uint32_t rounds = P.full_rounds + (P.tail_tiles ? 1u : 0u);
    for (uint32_t round = 0; round < rounds; round++) {
        if (P.sync && round > 0) {   // round barrier: every sweep starts together
            // A rendezvous in time only: no CTA reads what another wrote, so there is deliberately NO memory fence here -- a
            // fence would wait for this CTA's parity stores to be performed, and with the table in a peer GPU's memory
            // that is an NVLink drain per round (measured at 8 GPUs: 0.72 ms per step instead of 0.5).
            __syncthreads();
            if (threadIdx.x == 0) {
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
        const bool rev = P.serpentine && (round & 1);
        // full round: this CTA's tile, the whole sweep.  Shared last round: T sweeps of work in gridDim.x equal slices
        // (fixed point, so neighbours agree on the cut); a slice is a range of chunk groups of one tile, or the end of one
        // tile's sweep and the start of the next one's.
        uint32_t tile0 = round * gridDim.x + blockIdx.x, lo = 0, hi = tile0 < P.n_tiles ? HG_TAIL_K : 0;
        if (round >= P.full_rounds) {
            tile0 = P.full_rounds * gridDim.x;
            lo = (uint32_t)((uint64_t)blockIdx.x * P.tail_tiles * HG_TAIL_K / gridDim.x);
            hi = (uint32_t)((uint64_t)(blockIdx.x + 1) * P.tail_tiles * HG_TAIL_K / gridDim.x);
        }
        for (uint32_t t = lo / HG_TAIL_K; t * HG_TAIL_K < hi; t++) {
            const uint32_t a = max(lo, t * HG_TAIL_K) - t * HG_TAIL_K, b = min(hi, (t + 1) * HG_TAIL_K) - t * HG_TAIL_K;
            hg_sweep<VT, G, NV, NTAB, NB, XB, U, FULL>(P, T, tile0 + t, a, b, rev);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dispatch (per translation unit)
// ---------------------------------------------------------------------------------------------
template <typename KERN>
static int launch_hg(KERN kern, int smem, const HintParams &P, uint32_t grid, cudaStream_t st) {
    PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (P.sync) {
        // round barrier needs every CTA resident: cooperative launch (one 512-thread CTA per SM always fits)
        PM_CUDA(cudaMemsetAsync(P.sync, 0, sizeof(unsigned int), st));
        void *args[] = {(void *)&P};
        PM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(P.threads), args, (size_t)smem, st));
    } else {
        kern<<<grid, P.threads, smem, st>>>(P);
    }
    PM_CHECK_LAUNCH();
    count_launch();
    return PM_OK;
}
template <typename VT, int G, int NV, int NB, int XB>
static int launch_hintgen_t(const HintParams &P, uint32_t grid, cudaStream_t st) {
    constexpr int U = (G >= 2) ? 2 : 1;
    constexpr bool kFourTables = sizeof(VT) == 16 && NB == 2;  // instantiated only where it is ever selected
    // four T-tables win while the PRF dominates, i.e. rows up to 640 B; one table + PRMT rotations wins for 896 B rows
    // (measured on B200); the hg_ntab knob forces
    const int forced = tune(T_HG_NTAB);
    const int ntab = forced ? forced : (NV * G <= 40 ? 4 : 1);
    const bool full = (P.evx == (uint32_t)(NV * G));
    if (kFourTables && ntab == 4) {
        if (full) return launch_hg(hintgen_kernel<VT, G, NV, kFourTables ? 4 : 1, NB, XB, U, true>, aes_tab_words<4>() * 4, P, grid, st);
        return launch_hg(hintgen_kernel<VT, G, NV, kFourTables ? 4 : 1, NB, XB, U, false>, aes_tab_words<4>() * 4, P, grid, st);
    }
    if (full) return launch_hg(hintgen_kernel<VT, G, NV, 1, NB, XB, U, true>, aes_tab_words<1>() * 4, P, grid, st);
    return launch_hg(hintgen_kernel<VT, G, NV, 1, NB, XB, U, false>, aes_tab_words<1>() * 4, P, grid, st);
}
template <typename VT, int G, int NB, int XB>
static int launch_hintgen_nv(int nv, const HintParams &P, uint32_t grid, cudaStream_t st) {
    switch (nv) {
    case 1: return launch_hintgen_t<VT, G, 1, NB, XB>(P, grid, st);
    case 2: return launch_hintgen_t<VT, G, 2, NB, XB>(P, grid, st);
    case 3: return launch_hintgen_t<VT, G, 3, NB, XB>(P, grid, st);
    case 4: return launch_hintgen_t<VT, G, 4, NB, XB>(P, grid, st);
    case 5: return launch_hintgen_t<VT, G, 5, NB, XB>(P, grid, st);
    case 6: return launch_hintgen_t<VT, G, 6, NB, XB>(P, grid, st);
    case 7: return launch_hintgen_t<VT, G, 7, NB, XB>(P, grid, st);
    case 8: return launch_hintgen_t<VT, G, 8, NB, XB>(P, grid, st);
    }
    return set_error(PM_ERR_UNSUPPORTED, "hintgen: entry too wide (nv=%d)", nv);
}
template <typename VT, int NB, int XB>
static int launch_hintgen_g(const HintParams &P, uint32_t grid, cudaStream_t st) {
    int G, nv;
    hg_shape(P.evx, &G, &nv);
    switch (G) {
    case 8: return launch_hintgen_nv<VT, 8, NB, XB>(nv, P, grid, st);
    case 4: return nv == 1 ? launch_hintgen_t<VT, 4, 1, NB, XB>(P, grid, st) : launch_hintgen_t<VT, 4, 2, NB, XB>(P, grid, st);
    case 2: return nv == 1 ? launch_hintgen_t<VT, 2, 1, NB, XB>(P, grid, st) : launch_hintgen_t<VT, 2, 2, NB, XB>(P, grid, st);
    default: return launch_hintgen_t<VT, 1, 1, NB, XB>(P, grid, st);
    }
}

}  // namespace pm
