// bk_codegen.h -- generated marching kernels (bk_codegen.cu): tap list -> CUDA source -> NVRTC -> launch
#pragma once
#include <string>
#include <vector>
#include "bk_common.h"

namespace bk {

struct GenTap {
  int di, dj, dk;
  double c;
};
struct GenStencil;

// nullptr (and a reason in *why) when the taps are out of the generator's range or NVRTC is unavailable / fails
GenStencil *gen_create(const std::vector<GenTap> &taps, const bk_pointwise_t &pre, const bk_pointwise_t &post, std::string *why);
void gen_destroy(GenStencil *g);
const std::string &gen_source(const GenStencil *g);
size_t gen_cubin_bytes(const GenStencil *g);
// BK_EUNSUPPORTED for storage layouts the marching kernels cannot take (unaligned pointers, odd steps)
int gen_launch(GenStencil *g, const bk_field_t &f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
               const unsigned *hi, cudaStream_t s, int part, const unsigned *ready_lo, const unsigned *ready_hi);

// bk_stencil_tiled.cu
int launch_generated(const void *kernel, const GenGeom &g, const void *coef, const bk_field_t &f, const unsigned *grid,
                     const unsigned *gdims, const unsigned *lo, const unsigned *hi, cudaStream_t s, int part,
                     const unsigned *ready_lo, const unsigned *ready_hi);
size_t tiled_args_bytes();

}  // namespace bk
