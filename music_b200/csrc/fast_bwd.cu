// fast_bwd.cu - bf16 tensor-core backward (placeholder until the kernels land).
#include "fast_kernels.cuh"
#include "fast_layout.cuh"

namespace wn {

int build_bwd_maps(const Model&, const PackLayout&, const WsLayout&, int, int, const uint8_t*, uint8_t*,
                   const std::vector<CUtensorMap>&, BwdMaps*) {
  return WN_OK;
}

int fast_backward_impl(Model&, const BwdMaps&, int, int, const float*, const int64_t*, const void*, void*, float*, float*,
                       cudaStream_t) {
  set_error("wn_backward: bf16 backward kernels not built yet");
  return WN_ERR_UNSUPPORTED;
}

}  // namespace wn
