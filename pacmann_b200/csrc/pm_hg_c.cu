// hintgen_kernel instantiations: the general variants: any chunk_size / set_size, and 8-byte rows (odd entry_u64).  See pm_hintgen.cuh.
#include <stdint.h>
namespace pm {
__constant__ uint32_t c_te0_hg_c[256];
}
#define PM_HG_TE0 c_te0_hg_c
#include "pm_hintgen.cuh"

namespace pm {
int hg_upload_tables_c(const uint32_t te0[256]) {
    PM_CUDA(cudaMemcpyToSymbol(c_te0_hg_c, te0, 256 * sizeof(uint32_t)));
    return PM_OK;
}
int hg_launch_wide_xb4(const HintParams &P, uint32_t grid, cudaStream_t st) { return launch_hintgen_g<uint4, 4, 4>(P, grid, st); }
int hg_launch_narrow_xb2(const HintParams &P, uint32_t grid, cudaStream_t st) { return launch_hintgen_g<uint2, 2, 2>(P, grid, st); }
int hg_launch_narrow_xb4(const HintParams &P, uint32_t grid, cudaStream_t st) { return launch_hintgen_g<uint2, 4, 4>(P, grid, st); }
}  // namespace pm
