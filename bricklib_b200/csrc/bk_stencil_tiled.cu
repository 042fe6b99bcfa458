// bk_stencil_tiled.cu -- placeholder until the marching kernel lands: report "unsupported" so the dispatcher in
// bk_stencil.cu falls through to the per-brick family (a CUDA kernel, never a CPU path).
#include "bk_common.h"
namespace bk {
int launch_tiled(int, const bk_field_t &, const unsigned *, const unsigned *, const unsigned *, const unsigned *,
                 const double *, cudaStream_t) {
  return BK_EUNSUPPORTED;
}
}  // namespace bk
