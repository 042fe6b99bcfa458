// drivers/single.cpp -- the single-GPU self-validating benchmark: same flow and stdout lines as the reference's
// single/cuda.cpp -> d3pt7cu() (stencils/3axis.cu:77-198), with the stencil selectable at run time.
//
//   init_grid -> random array -> copyToBrick (host, in/out interleaved in ONE storage: step = 2*512, 3axis.cu:120-123)
//   -> movBrickInfo / movBrickStorage -> CPU array sweep = golden -> GPU sweeps timed with cutime_func
//   -> copy back -> compareBrick -> "result match" or std::runtime_error("result mismatch!").
//
// `-s cond` is the reference's second single-GPU case, d3condcu() (stencils/3axis.cu:242-330, stencils/cond.py): the 7-point
// star with every value read clamped at zero and the sum returned as its absolute value, run through the tap-table
// kernel (BrickStencilDef + pointwise clamps).  Inputs are shifted to [-0.5,0.5) and two coefficients negated so that both
// clamps matter (with the reference's all-positive data they never trigger).
//
// usage: single [-n cells_per_axis=512] [-s 7pt|mpi7pt|mpi13pt|mpi25pt|mpi125pt|cond] [-r launches=100]
#include <unistd.h>
#include <memory>
#include "common.h"

int main(int argc, char **argv) {
  long N = 512;
  std::string sname = "7pt";
  int reps = 100, c;
  while ((c = getopt(argc, argv, "n:s:r:h")) != -1) {
    if (c == 'n') N = atol(optarg);
    else if (c == 's') sname = optarg;
    else if (c == 'r') reps = atoi(optarg);
    else {
      std::cout << "usage: single [-n N] [-s stencil] [-r reps]" << std::endl;
      return c == 'h' ? 0 : 1;
    }
  }
  static const StencilDef kCond = {"cond", "stencils/cond.py", BK_ST_7PT, 1, 8, 7};
  const bool cond = sname == "cond";
  const StencilDef *st = cond ? &kCond : find_stencil(sname);
  if (!st || N % TILE) {
    std::cerr << "unknown stencil or N not a multiple of " << TILE << std::endl;
    return 1;
  }
  const long STRIDE = N + 2 * (GZ + PADDING), STRIDEG = N + 2 * GZ, STRIDEB = STRIDEG / TILE, NB = N / TILE, GB = GZ / TILE;

  // single/cpu.cpp:11-17: 129 random coefficients (the 7pt stencil uses the first 7)
  std::vector<bElem> coeff(129);
  {
    std::mt19937_64 rng(42);
    std::uniform_real_distribution<bElem> d(0, 1);
    for (auto &x : coeff) x = d(rng);
  }
  if (cond) coeff[2] = -coeff[2], coeff[5] = -coeff[5];
  bkCheck(bk_set_device(0));

  unsigned *grid_ptr;
  BrickInfo<3> bInfo = init_grid<3>(grid_ptr, {STRIDEB, STRIDEB, STRIDEB});
  bElem *in_ptr = randomArray({STRIDE, STRIDE, STRIDE});
  bElem *out_ptr = zeroArray({STRIDE, STRIDE, STRIDE});
  if (cond)
    for (long p = 0; p < STRIDE * STRIDE * STRIDE; ++p) in_ptr[p] -= 0.5;

  const unsigned bSize = cal_size<BDIM>::value;
  BrickStorage bStorage = bInfo.allocate(bSize * 2);
  std::memset(bStorage.dat.get(), 0, (size_t) bStorage.chunks * bStorage.step * sizeof(bElem));
  Brick3D bIn(&bInfo, bStorage, 0), bOut(&bInfo, bStorage, bSize);
  copyToBrick<3>({STRIDEG, STRIDEG, STRIDEG}, {PADDING, PADDING, PADDING}, {0, 0, 0}, in_ptr, grid_ptr, bIn);

  BrickInfo<3> bInfo_dev = movBrickInfo(bInfo, brickMemcpyHostToDevice);
  BrickStorage bStorage_dev = movBrickStorage(bStorage, brickMemcpyHostToDevice);
  Brick3D bIn_dev(&bInfo_dev, bStorage_dev, 0), bOut_dev(&bInfo_dev, bStorage_dev, bSize);
  unsigned *grid_dev = nullptr;
  copyToDevice({STRIDEB, STRIDEB, STRIDEB}, grid_dev, grid_ptr);

  // golden: plain array sweep on the host (arr_func, 3axis.cu:172)
  const std::vector<Tap> taps = stencil_taps(st->id, coeff.data());
  const long lo[3] = {PADDING + GZ, PADDING + GZ, PADDING + GZ}, hi[3] = {lo[0] + N, lo[1] + N, lo[2] + N};
  const std::vector<long> astride = {1, STRIDE, STRIDE * STRIDE};
  auto arr_func = [&]() {
    if (cond) cpu_array_sweep_cond(taps, in_ptr, out_ptr, astride, lo, hi);
    else cpu_array_sweep(taps, in_ptr, out_ptr, astride, lo, hi);
  };
  std::unique_ptr<BrickStencilDef> cond_def;  // cond.py lowered to its taps + clamps (what `python -m bricklib_b200.dsl` emits)
  if (cond) {
    std::vector<bk_tap_t> ctaps;
    for (const Tap &t : taps) ctaps.push_back({t.di, t.dj, t.dk, t.c});
    cond_def.reset(new BrickStencilDef(ctaps, bk_pointwise_t{BK_OP_MAX, 0.0}, bk_pointwise_t{BK_OP_ABS, 0.0}));
  }

  const std::vector<long> gd = {STRIDEB, STRIDEB, STRIDEB}, blo = {GB, GB, GB}, bhi = {NB + GB, NB + GB, NB + GB};
  auto brick_func = [&]() {
    if (cond) cond_def->launch(grid_dev, gd, bIn_dev, bOut_dev, blo, bhi, nullptr, BK_KERNEL_BRICK);
    else brickStencil(st->id, grid_dev, gd, bIn_dev, bOut_dev, blo, bhi, coeff.data(), nullptr, BK_KERNEL_BRICK);
  };
  auto brick_func_trans = [&]() {
    if (cond) cond_def->launch(grid_dev, gd, bIn_dev, bOut_dev, blo, bhi, nullptr, BK_KERNEL_AUTO);
    else brickStencil(st->id, grid_dev, gd, bIn_dev, bOut_dev, blo, bhi, coeff.data(), nullptr, BK_KERNEL_AUTO);
  };

  std::cout << (cond ? "d3cond" : "d3pt") << st->points << " (" << st->script << ", N = " << N << ")" << std::endl;
  std::cout << "Arr: " << time_func(arr_func, 1.0) << " (host, " << omp_get_max_threads() << " threads; the golden)" << std::endl;
  std::cout << "Bri: " << cutime_func(brick_func, std::max(1, reps / 10)) << std::endl;
  const double t = cutime_func(brick_func_trans, reps);
  std::cout << "Trans: " << t << std::endl;
  std::cout << "perf " << (double) N * N * N * 1e-9 / t << " GStencil/s, " << 16.0 * N * N * N * 1e-9 / t << " GB/s algorithmic" << std::endl;

  // copy the whole storage back (3axis.cu:179) and compare with the golden array
  BrickStorage back = movBrickStorage(bStorage_dev, brickMemcpyDeviceToHost);
  Brick3D bOutHost(&bInfo, back, bSize);
  if (!compareBrick<3>({N, N, N}, {PADDING, PADDING, PADDING}, {GZ, GZ, GZ}, out_ptr, grid_ptr, bOutHost))
    throw std::runtime_error("result mismatch!");
  std::cout << "result match" << std::endl;

  freeBrickInfoDevice(bInfo_dev);
  bk_dev_free(grid_dev);
  free(bInfo.adj);
  free(grid_ptr);
  free(in_ptr);
  free(out_ptr);
  return 0;
}
