// bk_march_common.h -- what the hand-written marching kernels (bk_stencil_tiled.cu) and the GENERATED ones (bk_codegen.cu
// emits CUDA source per stencil script and compiles it with NVRTC) share: the launch descriptor and the PTX helpers.
// This file is also embedded verbatim into every generated translation unit (bk_march_common.inc), so it must compile
// under NVRTC without any host header: built-in types only.
#ifndef BK_MARCH_COMMON_H
#define BK_MARCH_COMMON_H

struct bk_field_dev {  // = bk_field_t (include/bricklib_b200.h), restated with built-in types
  const unsigned *adj;
  const double *in;
  unsigned long long in_step;
  double *out;
  unsigned long long out_step;
};

struct TiledArgs {
  const double *in;
  double *out;
  unsigned long long in_step, out_step;  // elements between consecutive bricks
  const unsigned *grid;
  int gx, gy, gz;  // grid extents in bricks
  int lo[3], hi[3];
  int ntx;  // tiles along i
  int kl;   // brick layers per k segment
  int kh, kt;  // split launches: thin head / tail segments (layers) that keep the ghost-dependent CTAs few; else 0
  const bk_field_dev *multi;  // strong-scaling launch: per-subdomain fields (device array), subdomain = blockIdx.z
  // CTA enumeration: blockIdx.x runs through up to 6 boxes of the (tile i, tile j, k segment) space in order.  A plain
  // launch has one box.  A split launch (bk_stencil_apply_part) runs either the CTAs whose whole read footprint lies
  // inside the caller's "ready" brick box, or all the others -- the first part overlaps the ghost exchange, the second
  // is enqueued behind it, and both use the tile decomposition of the full box (no thin slab launches).
  int nbox;
  struct Box {
    int lo[3], dim[3], first;
  } box[6];
};

// Ghost bricks read straight out of the neighbours' storages (the exchange inside the sweep): brick id ghost_lo + g of the
// INPUT field is not read from in + id*step but from remap[g] -- the address of the peer's skin brick that this ghost
// brick mirrors (a CUDA-IPC / peer mapping, or the own storage for a self-neighbour).  Third argument of the *_remote
// kernels only; the other kernels do not know it exists.
struct RemoteArgs {
  const double *const *remap;
  unsigned ghost_lo, ghost_n;
};

// k range of segment `q`: [head of kh layers] [uniform segments of kl layers] [tail of kt layers]
__host__ __device__ __forceinline__ void seg_range(const TiledArgs &a, int q, int &kb0, int &nl) {
  const int nz = a.hi[2] - a.lo[2], mid = nz - a.kh - a.kt;
  if (a.kh > 0) {
    if (q == 0) {
      kb0 = a.lo[2], nl = a.kh;
      return;
    }
    --q;
  }
  if (q * a.kl < mid) {
    kb0 = a.lo[2] + a.kh + q * a.kl;
    nl = min(a.kl, mid - q * a.kl);
  } else {
    kb0 = a.hi[2] - a.kt, nl = a.kt;
  }
}

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

#endif  // BK_MARCH_COMMON_H
