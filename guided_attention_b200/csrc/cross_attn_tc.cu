// K1 / K2, tcgen05 variant (placeholder until the tensor-core kernels land).
#include "ga_common.cuh"
namespace ga {
namespace tc {
bool supports_fwd(int, int, int, int, bool) { return false; }
bool supports_bwd(int, int, int, int, bool) { return false; }
int fwd(const void*, const void*, const void*, void*, float*, float*, int, int, int, int, int, float, int, cudaStream_t) {
  return fail(GA_ERR_UNSUPPORTED, "tcgen05 forward not built");
}
int bwd(const void*, const void*, const void*, const float*, const void*, const float*, int64_t, void*, int, int, int,
        int, int, float, int, cudaStream_t) {
  return fail(GA_ERR_UNSUPPORTED, "tcgen05 backward not built");
}
}  // namespace tc
}  // namespace ga
