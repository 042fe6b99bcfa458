// drivers/scripts.cpp -- stencil scripts through the brick(...) statement, the way the reference's drivers use them.
//
// Each "kernel" below is the reference's pattern (stencils/3axis.cu:28-37, weak/main.cu:35-43): a function whose body is
// ONE brick(script, vec, dims, fold, b) statement; the script's grids (in/out, bIn/bOut, a/b) and constants (MPI_ALPHA,
// coeff[n], ...) are free variables resolved in the function's scope.  The build runs `python -m bricklib_b200.vecscatter`
// over this file (drivers/Makefile), as the reference runs codegen/vecscatter: the statement is replaced by code.  Here
// that code lowers the script to taps and launches the compiled stencil over the BrickLaunch `b` -- built-in marching
// kernels for stars and the symmetric cube, a kernel GENERATED for the tap pattern (NVRTC) for anything else.
// Every result is checked against a plain host array sweep with independently written taps.
//
// usage: scripts [-n cells_per_axis=64] [-r launches=20]
#include <unistd.h>
#include <functional>
#include <memory>
#include "common.h"
#include "vecscatter.h"

#define VSVEC "CUDA"
/* stencils/fake.h:11-33: the constants the mpi*.py scripts name */
#define MPI_ALPHA 0.4
#define MPI_BETA 0.1
#define MPI_C0 0.1
#define MPI_C1 0.04
#define MPI_C2 0.03
#define MPI_C3 0.01
#define MPI_C4 0.006
#define MPI_C5 0.004
#define MPI_C6 0.005
#define MPI_C7 0.002
#define MPI_C8 0.003
#define MPI_C9 0.001

// the scripts' grid names are the parameter names, exactly as in the reference's kernels
void k_mpi7pt(const BrickLaunch &b, Brick3D &in, Brick3D &out) { brick("../bricklib_b200/stencils/mpi7pt.py", VSVEC, (BDIM), (VFOLD), b); }
void k_mpi125pt(const BrickLaunch &b, Brick3D &in, Brick3D &out) { brick("../bricklib_b200/stencils/mpi125pt.py", VSVEC, (BDIM), (VFOLD), b); }
void k_7pt(const BrickLaunch &b, Brick3D &bIn, Brick3D &bOut, const bElem *coeff) { brick("../bricklib_b200/stencils/7pt.py", VSVEC, (BDIM), (VFOLD), b); }
void k_cond(const BrickLaunch &b, Brick3D &bIn, Brick3D &bOut, const bElem *coeff) { brick("../bricklib_b200/stencils/cond.py", VSVEC, (BDIM), (VFOLD), b); }
void k_box27_skewed(const BrickLaunch &L, Brick3D &a, Brick3D &b) { brick("../tests/stencil_scripts/box27_skewed.py", VSVEC, (BDIM), (VFOLD), L); }
void k_upwind(const BrickLaunch &L, Brick3D &u, Brick3D &v, double W) { brick("../tests/stencil_scripts/upwind.py", VSVEC, (BDIM), (VFOLD), L); }

int main(int argc, char **argv) {
  long N = 64;
  int reps = 20, c;
  while ((c = getopt(argc, argv, "n:r:h")) != -1) {
    if (c == 'n') N = atol(optarg);
    else if (c == 'r') reps = atoi(optarg);
    else {
      std::cout << "usage: scripts [-n N] [-r reps]" << std::endl;
      return c == 'h' ? 0 : 1;
    }
  }
  if (N % TILE) return 1;
  const long STRIDE = N + 2 * (GZ + PADDING), STRIDEG = N + 2 * GZ, STRIDEB = STRIDEG / TILE, NB = N / TILE, GB = GZ / TILE;
  std::vector<bElem> coeff(7);
  {
    std::mt19937_64 rng(42);
    std::uniform_real_distribution<bElem> d(0, 1);
    for (auto &x : coeff) x = d(rng);
  }
  bkCheck(bk_set_device(0));
  unsigned *grid_ptr;
  BrickInfo<3> bInfo = init_grid<3>(grid_ptr, {STRIDEB, STRIDEB, STRIDEB});
  bElem *in_ptr = randomArray({STRIDE, STRIDE, STRIDE});
  for (long p = 0; p < STRIDE * STRIDE * STRIDE; ++p) in_ptr[p] -= 0.25;  // some negative values: cond.py's clamps matter
  bElem *out_ptr = zeroArray({STRIDE, STRIDE, STRIDE});
  const unsigned bSize = cal_size<BDIM>::value;
  BrickStorage bStorage = bInfo.allocate(bSize * 2);
  std::memset(bStorage.dat.get(), 0, (size_t) bStorage.chunks * bStorage.step * sizeof(bElem));
  Brick3D bIn(&bInfo, bStorage, 0);
  copyToBrick<3>({STRIDEG, STRIDEG, STRIDEG}, {PADDING, PADDING, PADDING}, {0, 0, 0}, in_ptr, grid_ptr, bIn);
  BrickInfo<3> bInfo_dev = movBrickInfo(bInfo, brickMemcpyHostToDevice);
  BrickStorage bStorage_dev = movBrickStorage(bStorage, brickMemcpyHostToDevice);
  Brick3D bIn_dev(&bInfo_dev, bStorage_dev, 0), bOut_dev(&bInfo_dev, bStorage_dev, bSize);
  unsigned *grid_dev = nullptr;
  copyToDevice({STRIDEB, STRIDEB, STRIDEB}, grid_dev, grid_ptr);
  const BrickLaunch box(grid_dev, {STRIDEB, STRIDEB, STRIDEB}, {GB, GB, GB}, {NB + GB, NB + GB, NB + GB});

  const long lo[3] = {PADDING + GZ, PADDING + GZ, PADDING + GZ}, hi[3] = {lo[0] + N, lo[1] + N, lo[2] + N};
  const std::vector<long> astride = {1, STRIDE, STRIDE * STRIDE};
  int failures = 0;
  auto check = [&](const char *name, const std::vector<Tap> &taps, bool clamps, const std::function<void()> &launch) {
    if (clamps) cpu_array_sweep_cond(taps, in_ptr, out_ptr, astride, lo, hi);
    else cpu_array_sweep(taps, in_ptr, out_ptr, astride, lo, hi);
    const double t = cutime_func(launch, reps);
    BrickStorage back = movBrickStorage(bStorage_dev, brickMemcpyDeviceToHost);
    Brick3D bOutHost(&bInfo, back, bSize);
    const bool ok = compareBrick<3>({N, N, N}, {PADDING, PADDING, PADDING}, {GZ, GZ, GZ}, out_ptr, grid_ptr, bOutHost);
    std::cout << name << ": " << t << " s per sweep, " << (double) N * N * N * 1e-9 / t << " GStencil/s, "
              << (ok ? "result match" : "result mismatch!") << std::endl;
    failures += !ok;
  };

  check("mpi7pt.py", stencil_taps(BK_ST_MPI7PT, nullptr), false, [&] { k_mpi7pt(box, bIn_dev, bOut_dev); });
  check("mpi125pt.py", stencil_taps(BK_ST_MPI125PT, nullptr), false, [&] { k_mpi125pt(box, bIn_dev, bOut_dev); });
  check("7pt.py", stencil_taps(BK_ST_7PT, coeff.data()), false, [&] { k_7pt(box, bIn_dev, bOut_dev, coeff.data()); });
  std::vector<bElem> ccoef = coeff;
  ccoef[2] = -ccoef[2], ccoef[5] = -ccoef[5];
  check("cond.py", stencil_taps(BK_ST_7PT, ccoef.data()), true, [&] { k_cond(box, bIn_dev, bOut_dev, ccoef.data()); });
  {
    std::vector<Tap> t;  // box27_skewed.py: weight (n+1)/100 for the n-th offset of product((-1,0,1), repeat=3), i slowest
    int n = 0;
    for (int di = -1; di <= 1; ++di)
      for (int dj = -1; dj <= 1; ++dj)
        for (int dk = -1; dk <= 1; ++dk) t.push_back({dk, dj, di, (++n) / 100.0});
    check("box27_skewed.py", t, false, [&] { k_box27_skewed(box, bIn_dev, bOut_dev); });
  }
  {
    const double W = 0.3;  // upwind.py
    const std::vector<Tap> t = {{0, 0, 0, 0.5}, {0, 0, -1, -W}, {0, 1, -2, 2 * W}, {3, -1, 1, 0.25}, {-3, 0, 0, -0.125},
                                {0, 2, 0, 0.25}, {0, -2, 0, -0.25}};
    check("upwind.py", t, false, [&] { k_upwind(box, bIn_dev, bOut_dev, W); });
  }
  freeBrickInfoDevice(bInfo_dev);
  bk_dev_free(grid_dev);
  return failures ? 2 : 0;
}
