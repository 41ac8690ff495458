// hintgen_kernel instantiations: 16-byte rows, chunk_size <= 65536, set_size <= 65536.  See pm_hintgen.cuh.
#include <stdint.h>
namespace pm {
__constant__ uint32_t c_te0_hg_b[256];
}
#define PM_HG_TE0 c_te0_hg_b
#include "pm_hintgen.cuh"

namespace pm {
int hg_upload_tables_b(const uint32_t te0[256]) {
    PM_CUDA(cudaMemcpyToSymbol(c_te0_hg_b, te0, 256 * sizeof(uint32_t)));
    return PM_OK;
}
int hg_launch_wide_xb2(const HintParams &P, uint32_t grid, cudaStream_t st) { return launch_hintgen_g<uint4, 2, 2>(P, grid, st); }
}  // namespace pm
