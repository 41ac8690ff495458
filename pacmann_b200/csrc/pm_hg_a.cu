// hintgen_kernel instantiations: 16-byte rows, chunk_size <= 65536, set_size <= 256 (PRF rounds 1-2 hoisted for one varying chunk-id byte).  See pm_hintgen.cuh.
#include <stdint.h>
namespace pm {
__constant__ uint32_t c_te0_hg_a[256];
}
#define PM_HG_TE0 c_te0_hg_a
#include "pm_hintgen.cuh"

namespace pm {
int hg_upload_tables_a(const uint32_t te0[256]) {
    PM_CUDA(cudaMemcpyToSymbol(c_te0_hg_a, te0, 256 * sizeof(uint32_t)));
    return PM_OK;
}
int hg_launch_wide_xb1(const HintParams &P, uint32_t grid, cudaStream_t st) { return launch_hintgen_g<uint4, 2, 1>(P, grid, st); }
}  // namespace pm
