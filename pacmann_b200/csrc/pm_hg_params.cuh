// Launch parameters of hintgen_kernel, shared by the host-side enqueue (pm_hintgen.cu) and the kernel translation units.
#pragma once
#include <stdint.h>

#include "pm_common.cuh"

namespace pm {

constexpr int HG_THREADS = 512;        // default CTA width (16 warps)
constexpr int HG_MAX_THREADS = 512;    // wider CTAs would cap the kernel below the 128 registers it needs (spills)
constexpr int HG_MAX_JOBS = 16;
constexpr uint32_t HG_TAIL_K = 1u << 12;   // fixed-point unit of the tail round's work split

struct HintJobDev {
    uint32_t rk[44];
    uint64_t row0, n_rows;
    uint64_t hint_begin, n_hints, n_primary, backup_group;
    const uint64_t *tags;
    const int32_t *skip;
    uint64_t *out;
    uint16_t *off;      // optional offset index output [n_hints][(set_size + 7) & ~7]
    uint32_t chunk_mask, chunk_shift, set_size, tile_begin;
};
struct HintParams {
    HintJobDev jobs[HG_MAX_JOBS];
    const void *db;
    uint32_t n_jobs, n_tiles;
    uint32_t threads;       // CTA width of this launch
    uint32_t ev, evx;       // vectors per row (incl. un-xored tail), vectors that are xored
    uint32_t full_rounds;   // rounds in which every CTA owns one whole tile
    uint32_t tail_tiles;    // tiles of the shared last round (0 = none; their outputs are zeroed before the launch)
    uint32_t serpentine;    // alternate the sweep direction between rounds
    unsigned int *sync;     // round barrier counter (cooperative launch) or nullptr
};
// lane-group width and vectors per lane for rows of `evx` xored vectors
static inline void hg_shape(uint32_t evx, int *G, int *nv) {
    const uint32_t x = evx ? evx : 1;
    *G = x >= 8 ? 8 : x >= 4 ? 4 : x >= 2 ? 2 : 1;
    *nv = (int)((x + *G - 1) / *G);
}
// one launcher per PRF variant (pm_hg_a.cu, pm_hg_b.cu, pm_hg_c.cu); NB = PRF output bytes kept, XB = chunk-id bytes that vary
int hg_launch_wide_xb1(const HintParams &P, uint32_t grid, cudaStream_t st);    // uint4 rows, chunk <= 65536, set <= 256
int hg_launch_wide_xb2(const HintParams &P, uint32_t grid, cudaStream_t st);    // uint4 rows, chunk <= 65536, set <= 65536
int hg_launch_wide_xb4(const HintParams &P, uint32_t grid, cudaStream_t st);    // uint4 rows, anything
int hg_launch_narrow_xb2(const HintParams &P, uint32_t grid, cudaStream_t st);  // uint2 rows (odd entry_u64)
int hg_launch_narrow_xb4(const HintParams &P, uint32_t grid, cudaStream_t st);
int hg_upload_tables_a(const uint32_t te0[256]);
int hg_upload_tables_b(const uint32_t te0[256]);
int hg_upload_tables_c(const uint32_t te0[256]);

}  // namespace pm
