// Fused GraphConv layer kernel (placeholder until the tensor-core kernel lands).
#include "common.cuh"

namespace kgcn {

bool fused_fwd_eligible(int64_t, int, int, int, int, const float*, const float*) { return false; }

int launch_graphconv_fused_fwd(const int32_t*, const int32_t*, const float*, int64_t, int, int, const float*, int,
                               const float*, const float*, int, int, float*, cudaStream_t) {
    return fail(KGCN_ERR_UNSUPPORTED, "fused GraphConv kernel not built");
}

}  // namespace kgcn
