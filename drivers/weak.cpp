// drivers/weak.cpp -- weak-scaling driver: one subdomain per GPU, ghost-zone exchange every ST_ITER sweeps.
//
// Same CLI letters and stdout lines as the reference's weak/main.cu + weak/args.cpp (-d/-s/-I/-b/-h; "Bri:", "calc",
// "move", "call", "wait", "  | MPI size (MB)", "  | MPI speed (GB/s)", "perf X GStencil/s"), same time loop
// (weak/main.cu:246-287: exchange, then ST_ITER sweeps ping-ponging between two storages over the whole grid
// including ghost bricks; the last sweep of a period skips the ghost shell like weak/main.cpp:209).
// What differs: ranks are host threads of ONE process, one per GPU (-g N; the reference gets N from mpirun), the
// periodic Cartesian "communicator" is a BrickComm, and the exchange is one pull kernel over NVLink peer pointers --
// ghost bricks never leave HBM (the reference stages 97.6 MB down and 103.8 MB up through the host per exchange,
// weak/main.cu:251-272).  The first sweep of a period runs on the inner bricks WHILE the pull is in flight.
//
// usage: weak [-s i,j,k | -d i,j,k] [-I periods] [-g gpus] [-S stencil] [-v] [-b]
//   -v  validate against a CPU sweep of the global periodic array (small domains / few periods only)
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include "array-mpi.h"
#include "common.h"

namespace {

void parseTuple(std::string s, std::vector<unsigned> &out) {
  for (auto &o : out) {
    size_t r = s.find(',');
    o = (unsigned) std::stoi(s.substr(0, r));
    s = r == std::string::npos ? "" : s.substr(r + 1);
    if (s.empty()) s = std::to_string(o);  // "-s 64" means 64,64,64
  }
}

void dims_create(int n, int *d) {  // MPI_Dims_create: balanced, non-increasing
  d[0] = d[1] = d[2] = 1;
  for (int p = 2; n > 1;) {
    if (n % p) {
      ++p;
      continue;
    }
    int *m = &d[0];
    for (int i = 1; i < 3; ++i)
      if (d[i] < *m) m = &d[i];
    *m *= p, n /= p;
  }
  std::sort(d, d + 3, [](int a, int b) { return a > b; });
}

struct Shared {
  int size = 1, cart[3] = {1, 1, 1}, iters = 100;
  std::vector<unsigned> dom = {64, 64, 64};
  const StencilDef *st = nullptr;
  bool validate = false;
  bool no_fuse = false;  // -F: one sweep per pass (no temporal blocking)
  bool no_array = false; // -A: skip the array-layout baseline ("Arr:" block)
  std::vector<bElem *> storage_ptr;  // rank -> device address of storage[0]
  std::vector<bElem *> array_ptr;    // rank -> device address of the input array of the Arr: baseline
  std::vector<double> calc, call, wait, total;
  std::vector<int> arr_match;        // rank -> brick result equals array result (compareBrick on the device)
  std::vector<bElem *> result;       // rank -> host copy of the interior after the run (validation)
  bElem *global_in = nullptr;
  std::atomic<int> failures{0};
};

void rank_main(int rank, Shared &S, Barrier &bar) {
  int ndev = 0;
  bkCheck(bk_device_count(&ndev));
  if (ndev == 0) throw std::runtime_error("no CUDA device: this build has no CPU path");
  const int dev = rank % ndev;
  bkCheck(bk_set_device(dev));
  BrickComm comm = BrickComm::cart(S.cart, rank);
  const std::vector<unsigned> &dom = S.dom;
  const StencilDef *st = S.st;

  std::vector<long> stride(3), strideg(3), strideb(3);
  for (int i = 0; i < 3; ++i) stride[i] = dom[i] + 2 * TILE + 2 * GZ, strideg[i] = dom[i] + 2 * TILE, strideb[i] = strideg[i] / TILE;

  BrickDecomp<3, BDIM> bDecomp(dom, GZ);
  populate(comm, bDecomp);
  bDecomp.initialize(skin3d_good);
  BrickInfo<3> bInfo = bDecomp.getBrickInfo();
  const unsigned bSize = cal_size<BDIM>::value;

  // adjacency symmetry self-check of every driver (weak/main.cu:102-109)
  for (long k = 1; k < strideb[2] - 1; ++k)
    for (long j = 1; j < strideb[1] - 1; ++j)
      for (long i = 1; i < strideb[0] - 1; ++i) {
        const unsigned l = bDecomp[k][j][i];
        for (int id = 0; id < 27; ++id)
          if (bInfo.adj[bInfo.adj[l][id]][26 - id] != l) throw std::runtime_error("err");
      }

  BrickInfo<3> bInfo_dev = movBrickInfo(bInfo, brickMemcpyHostToDevice);
  BrickStorage bStorage_dev = deviceBrickStorage(bInfo.nbricks, bSize), bStorageOut_dev = deviceBrickStorage(bInfo.nbricks, bSize);
  Brick3D bIn_dev(&bInfo_dev, bStorage_dev, 0), bOut_dev(&bInfo_dev, bStorageOut_dev, 0);
  unsigned *grid_dev = nullptr;
  copyToDevice(strideb, grid_dev, bDecomp.gridData());

  bElem *arr_in_dev = nullptr, *arr_out_dev = nullptr;  // the Arr: baseline's arrays (the input doubles as the bricks' source)
  // input: this rank's block of a global random periodic field (interior only; ghosts come from the first exchange)
  {
    bElem *in_ptr;
    if (S.validate) {
      in_ptr = zeroArray(stride);
      const long G[3] = {(long) dom[0] * S.cart[2], (long) dom[1] * S.cart[1], (long) dom[2] * S.cart[0]};
      const long o[3] = {(long) comm.coords[2] * dom[0], (long) comm.coords[1] * dom[1], (long) comm.coords[0] * dom[2]};
      for (long k = 0; k < dom[2]; ++k)
        for (long j = 0; j < dom[1]; ++j)
          std::memcpy(in_ptr + (PADDING + GZ) + (j + PADDING + GZ) * stride[0] + (k + PADDING + GZ) * stride[0] * stride[1],
                      S.global_in + o[0] + (o[1] + j) * G[0] + (o[2] + k) * G[0] * G[1], dom[0] * sizeof(bElem));
    } else {
      in_ptr = randomArray(stride, 0x5EED + rank);
    }
    copyToDevice(stride, arr_in_dev, in_ptr);
    copyToBrickDevice(strideg, {PADDING, PADDING, PADDING}, {0, 0, 0}, arr_in_dev, grid_dev, bIn_dev);
    bkCheck(bk_device_sync());
    free(in_ptr);
  }

  // ---- Arr: the array-layout baseline of the reference (weak/main.cu:161-213): exchangeArr + ST_ITER arr_kernel sweeps
  // over the whole ghost-inclusive region, on the same input, timed with the same protocol as the bricks below
  const std::vector<long> dom_l = {(long) dom[0], (long) dom[1], (long) dom[2]}, pad3 = {PADDING, PADDING, PADDING}, gz3 = {GZ, GZ, GZ};
  if (!S.no_array) {
    const size_t abytes = (size_t) stride[0] * stride[1] * stride[2] * sizeof(bElem);
    void *tmp = nullptr;
    bkCheck(bk_dev_alloc(&tmp, abytes));
    arr_out_dev = (bElem *) tmp;
    bkCheck(bk_dev_memset(arr_out_dev, 0, abytes, nullptr));
    S.array_ptr[rank] = arr_in_dev;
    bar.wait();
    for (auto &kv : bDecomp.rank_map)
      if (kv.second % ndev != dev) bkCheck(bk_peer_enable(kv.second % ndev));
    ArrayExchangeView aev(arr_in_dev, bDecomp.rank_map, S.array_ptr, dom_l, pad3, gz3);
    void *e0, *e1, *e2, *e3;
    for (void **e : {&e0, &e1, &e2, &e3}) bkCheck(bk_event_create(e));
    const std::vector<long> alo = pad3, ahi = {stride[0] - PADDING, stride[1] - PADDING, stride[2] - PADDING};
    double a_calc = 0, a_call = 0, a_wait = 0;
    auto arr_func = [&]() {
      bkCheck(bk_device_sync());
      double t0 = omp_get_wtime();
      bar.wait();  // every rank's previous sweeps are complete before anyone pulls
      a_wait += omp_get_wtime() - t0;
      bkCheck(bk_event_record(e0, nullptr));
      aev.exchange(nullptr);
      bkCheck(bk_event_record(e1, nullptr));
      bkCheck(bk_event_sync(e1));
      float ms = 0;
      bkCheck(bk_event_elapsed_ms(e0, e1, &ms));
      a_call += ms / 1e3;
      t0 = omp_get_wtime();
      bar.wait();  // every pull has finished before the second sweep overwrites an input array
      a_wait += omp_get_wtime() - t0;
      bkCheck(bk_event_record(e2, nullptr));
      for (int i = 0; i < st->st_iter / 2; ++i) {
        arrayStencil(st->id, arr_in_dev, arr_out_dev, stride, alo, ahi);
        arrayStencil(st->id, arr_out_dev, arr_in_dev, stride, alo, ahi);
      }
      bkCheck(bk_event_record(e3, nullptr));
      bkCheck(bk_event_sync(e3));
      bkCheck(bk_event_elapsed_ms(e2, e3, &ms));
      a_calc += ms / 1e3;
    };
    arr_func();
    a_calc = a_call = a_wait = 0;
    bar.wait();
    const double t_st = omp_get_wtime();
    for (int i = 0; i < S.iters; ++i) arr_func();
    bkCheck(bk_device_sync());
    bar.wait();
    const double per_period = (omp_get_wtime() - t_st) / S.iters;
    const int cnt = S.iters * st->st_iter;
    S.calc[rank] = a_calc / cnt, S.call[rank] = a_call / cnt, S.wait[rank] = a_wait / cnt, S.total[rank] = per_period / st->st_iter;
    bar.wait();
    if (rank == 0) {
      const double tsize = 2.0 * aev.bytes;  // sent + received
      mpi_stats calc_s = mpi_statistics(S.calc), call_s = mpi_statistics(S.call), wait_s = mpi_statistics(S.wait);
      mpi_stats tot_s = mpi_statistics(S.total);
      std::vector<double> spd(S.size), sz(S.size, tsize * 1e-6), zero(S.size, 0.0);
      for (int r = 0; r < S.size; ++r) spd[r] = tsize / 1e9 / std::max(1e-12, (S.call[r] + S.wait[r]) * st->st_iter);
      std::cout << "Arr: " << tot_s.max << std::endl;
      std::cout << "calc " << calc_s << std::endl;
      std::cout << "pack " << mpi_statistics(zero) << std::endl;
      std::cout << "move " << mpi_statistics(zero) << std::endl;
      std::cout << "call " << call_s << std::endl;
      std::cout << "wait " << wait_s << std::endl;
      std::cout << "  | MPI size (MB): " << mpi_statistics(sz) << std::endl;
      std::cout << "  | MPI speed (GB/s): " << mpi_statistics(spd) << std::endl;
      double tot_elems = (double) S.size * dom[0] * dom[1] * dom[2];
      std::cout << "perf " << tot_elems * 1.0e-9 / tot_s.max << " GStencil/s" << std::endl << std::endl;
    }
    bar.wait();
    for (void *e : {e0, e1, e2, e3}) bk_event_destroy(e);
  }

  // wire the neighbours: peer access + their storage addresses
  S.storage_ptr[rank] = bStorage_dev.dat.get();
  bar.wait();
  for (auto &kv : bDecomp.rank_map)
    if (kv.second % ndev != dev) bkCheck(bk_peer_enable(kv.second % ndev));  // idempotent; throws without P2P
  ExchangeView ev = bDecomp.exchangeView(bStorage_dev, S.storage_ptr);

  void *comm_stream, *evDone, *evX, *c0, *c1, *x0, *x1;
  bkCheck(bk_stream_create_priority(&comm_stream, 1));
  for (void **e : {&evDone, &evX, &c0, &c1, &x0, &x1}) bkCheck(bk_event_create(e));
  const std::vector<long> full_lo = {0, 0, 0}, full_hi = strideb;
  const long g = GZ / TILE;
  const std::vector<long> skip_lo = {g, g, g}, skip_hi = {strideb[0] - g, strideb[1] - g, strideb[2] - g};
  double calctime = 0, calltime = 0, waittime = 0;

  // time steps per pass: two where the fused kernel pays off (7-point), -F forces one sweep per pass
  int fuse = S.no_fuse ? 1 : bk_stencil_fused_steps(st->id);
  if (fuse < 1 || st->st_iter % fuse) fuse = 1;
  const int npass = st->st_iter / fuse;
  // pass p reads storage p%2 and writes the other one; the last pass of a period skips the ghost shell
  auto pass = [&](int p, int part, void *stream) -> bool {
    const bool last = p == npass - 1;
    auto &src = (p % 2) ? bOut_dev : bIn_dev;
    auto &dst = (p % 2) ? bIn_dev : bOut_dev;
    return brickAdvance(st->id, fuse, grid_dev, strideb, src, dst, last ? skip_lo : full_lo, last ? skip_hi : full_hi,
                        skip_lo, skip_hi, part, nullptr, stream);
  };
  // thin k segments for the ghost-dependent layers of the split pass: pays off when ghost ranges cross NVLink and the
  // radius is <= 2 (N = 8 sweep, profiles/r01c_multi_gpu.md)
  const int thin = (S.size > 1 && st->radius <= 2) ? BK_PART_THIN : 0;
  auto brick_func = [&]() {
    // every rank's previous sweeps are complete before anyone pulls
    bkCheck(bk_event_record(evDone, nullptr));
    bkCheck(bk_event_sync(evDone));
    double t0 = omp_get_wtime();
    bar.wait();
    waittime += omp_get_wtime() - t0;
    t0 = omp_get_wtime();
    bkCheck(bk_event_record(x0, comm_stream));
    ev.exchange(comm_stream);
    bkCheck(bk_event_record(evX, comm_stream));
    bkCheck(bk_event_record(c0, nullptr));
    // pass 0 in two launches over the same tiles: CTAs that read only my own bricks overlap the pull (compute
    // stream), the CTAs that touch the ghost shell follow the pull on the high-priority exchange stream
    if (pass(0, BK_PART_READY | thin, nullptr)) {
      pass(0, BK_PART_REST | thin, comm_stream);
      bkCheck(bk_event_record(evX, comm_stream));
      bkCheck(bk_stream_wait_event(nullptr, evX));
    } else {
      if (fuse != 1) throw std::runtime_error("fused pass unavailable for this storage layout: rerun with -F");
      bkCheck(bk_stream_wait_event(nullptr, evX));
      brickStencil(st->id, grid_dev, strideb, bIn_dev, bOut_dev, npass == 1 ? skip_lo : full_lo,
                   npass == 1 ? skip_hi : full_hi, nullptr, nullptr);
    }
    calltime += omp_get_wtime() - t0;
    // pass 1 overwrites storage 0, whose skin the neighbours are pulling: wait until every pull has finished
    bkCheck(bk_event_sync(evX));
    t0 = omp_get_wtime();
    bar.wait();
    waittime += omp_get_wtime() - t0;
    for (int p = 1; p < npass; ++p)
      if (!pass(p, BK_PART_ALL, nullptr)) throw std::runtime_error("pass unavailable");
    bkCheck(bk_event_record(c1, nullptr));
    bkCheck(bk_event_sync(c1));
    float ms = 0;
    bkCheck(bk_event_elapsed_ms(c0, c1, &ms));
    calctime += ms / 1e3;
  };

  brick_func();  // warm-up (time_mpi, stencils/fake.h:392-404)
  calctime = calltime = waittime = 0;
  bar.wait();
  const double st_t = omp_get_wtime();
  for (int i = 0; i < S.iters; ++i) brick_func();
  bkCheck(bk_device_sync());
  bar.wait();
  const double per_period = (omp_get_wtime() - st_t) / S.iters;
  const int cnt = S.iters * st->st_iter;
  S.calc[rank] = calctime / cnt, S.call[rank] = calltime / cnt, S.wait[rank] = waittime / cnt;
  S.total[rank] = per_period / st->st_iter;

  if (S.validate) {  // interior of the final field (storage 0 after an even number of sweeps) back to the host
    bElem *arr_dev = nullptr, *zero = zeroArray(stride);
    copyToDevice(stride, arr_dev, zero);
    copyFromBrickDevice({(long) dom[0], (long) dom[1], (long) dom[2]}, {PADDING, PADDING, PADDING}, {GZ, GZ, GZ}, arr_dev,
                        grid_dev, (npass % 2) ? bOut_dev : bIn_dev);
    copyFromDevice(stride, zero, arr_dev);
    bk_dev_free(arr_dev);
    S.result[rank] = zero;
  }
  if (!S.no_array) {  // the reference's closing check: the brick result equals the array result (weak/main.cu:325-327)
    S.arr_match[rank] = compareBrickDevice(dom_l, pad3, gz3, arr_in_dev, grid_dev, (npass % 2) ? bOut_dev : bIn_dev, BRICK_TOLERANCE);
    bk_dev_free(arr_out_dev);
  }
  bk_dev_free(arr_in_dev);
  bar.wait();
  if (rank == 0) {
    const size_t tsize = bDecomp.exchangeSize() * bSize * sizeof(bElem) * 2;  // sent + received, as weak/main.cpp:220-222
    mpi_stats calc_s = mpi_statistics(S.calc), call_s = mpi_statistics(S.call), wait_s = mpi_statistics(S.wait);
    mpi_stats tot_s = mpi_statistics(S.total);
    std::vector<double> spd(S.size), sz(S.size, tsize * 1e-6), zero(S.size, 0.0);
    for (int r = 0; r < S.size; ++r) spd[r] = tsize / 1e9 / std::max(1e-12, (S.call[r] + S.wait[r]) * st->st_iter);
    std::cout << "Bri: " << tot_s.max << std::endl;
    std::cout << "calc " << calc_s << std::endl;
    std::cout << "move " << mpi_statistics(zero) << std::endl;
    std::cout << "call " << call_s << std::endl;
    std::cout << "wait " << wait_s << std::endl;
    std::cout << "  | MPI size (MB): " << mpi_statistics(sz) << std::endl;
    std::cout << "  | MPI speed (GB/s): " << mpi_statistics(spd) << std::endl;
    double tot_elems = (double) S.size * dom[0] * dom[1] * dom[2];
    std::cout << "perf " << tot_elems * 1.0e-9 / tot_s.max << " GStencil/s" << std::endl;
    std::cout << "Total of " << bDecomp.ghost.size() << " parts" << std::endl;
    if (!S.no_array) {
      bool all = true;
      for (int m : S.arr_match) all = all && m;
      std::cout << (all ? "Arr == Bri: result match" : "Arr == Bri: result mismatch!") << std::endl;
      if (!all) S.failures++;
    }
  }
  freeBrickInfoDevice(bInfo_dev);
  bk_dev_free(grid_dev);
  bk_stream_destroy(comm_stream);
}

}  // namespace

int main(int argc, char **argv) {
  Shared S;
  std::string sname = "mpi7pt";
  int c, sel = 0;
  bool bin = false;
  if (const char *e = getenv("BRICK_RANKS")) S.size = atoi(e);
  while ((c = getopt(argc, argv, "d:s:I:g:S:vbhFA")) != -1) switch (c) {
      case 'b': bin = true; break;
      case 'd': parseTuple(optarg, S.dom), sel = sel ? -2 : 1; break;
      case 's': parseTuple(optarg, S.dom), sel = sel ? -2 : 2; break;
      case 'I': S.iters = std::stoi(optarg); break;
      case 'g': S.size = std::stoi(optarg); break;
      case 'S': sname = optarg; break;
      case 'v': S.validate = true; break;
      case 'F': S.no_fuse = true; break;
      case 'A': S.no_array = true; break;
      default:
        printf("Program options\n  -h: help\n  -b: process grid of powers of two\n  -d i,j,k: overall domain size\n"
               "  -s i,j,k: per-GPU domain size\n  -I n: exchange periods (default 100)\n  -g n: GPUs = ranks (default 1)\n"
               "  -S name: 7pt mpi7pt mpi13pt mpi25pt mpi125pt\n  -v: validate against a CPU sweep\n  -F: one sweep per pass\n"
               "  -A: skip the array-layout baseline (Arr: block)\n");
        return 0;
    }
  if (sel == -2) {
    printf("Contradicting options\n");
    return 0;
  }
  S.st = find_stencil(sname);
  if (!S.st || S.st->id == BK_ST_7PT) {
    std::cerr << "stencil must be one of mpi7pt mpi13pt mpi25pt mpi125pt" << std::endl;
    return 1;
  }
  if (bin) {
    int b = 0, s = 1;
    while (s * 2 <= S.size) S.cart[b++ % 3] *= 2, s *= 2;
    S.size = s;
  } else {
    dims_create(S.size, S.cart);
  }
  if (sel == 1)  // overall size: split evenly (must divide)
    for (int i = 0; i < 3; ++i) S.dom[i] = S.dom[i] / S.cart[2 - i] / TILE * TILE;
  size_t tot = (size_t) S.size * S.dom[0] * S.dom[1] * S.dom[2];
  std::cout << "Pagesize " << sysconf(_SC_PAGESIZE) << "; MPI Size " << S.size << " * OpenMP threads " << omp_get_max_threads() << std::endl;
  std::cout << "Domain size of " << tot << " split among" << std::endl;
  std::cout << "A total of " << S.size << " processes " << S.cart[0] << "x" << S.cart[1] << "x" << S.cart[2] << std::endl;
  std::cout << "d3pt" << S.st->points << " MPI decomp (" << S.st->script << ", one rank per GPU, NVLink pull exchange)" << std::endl;

  S.storage_ptr.assign(S.size, nullptr);
  S.array_ptr.assign(S.size, nullptr);
  S.arr_match.assign(S.size, 1);
  S.calc.assign(S.size, 0), S.call = S.wait = S.total = S.calc;
  S.result.assign(S.size, nullptr);
  const long G[3] = {(long) S.dom[0] * S.cart[2], (long) S.dom[1] * S.cart[1], (long) S.dom[2] * S.cart[0]};
  if (S.validate) S.global_in = randomArray({G[0], G[1], G[2]});

  Barrier bar(S.size);
  std::vector<std::thread> th;
  std::atomic<bool> failed{false};
  for (int r = 0; r < S.size; ++r)
    th.emplace_back([&, r] {
      try {
        rank_main(r, S, bar);
      } catch (const std::exception &e) {
        std::cerr << "rank " << r << ": " << e.what() << std::endl;
        failed = true;
        exit(EXIT_FAILURE);
      }
    });
  for (auto &t : th) t.join();

  if (S.validate) {  // golden: the global periodic array advanced (I+1)*ST_ITER steps on the host
    const int steps = (S.iters + 1) * S.st->st_iter;
    cpu_periodic_steps(S.global_in, G, stencil_taps(S.st->id, nullptr), S.st->radius, steps);
    const bElem *cur = S.global_in;
    long bad = 0;
    double worst = 0;
    for (int r = 0; r < S.size; ++r) {
      BrickComm cm = BrickComm::cart(S.cart, r);
      const long o[3] = {(long) cm.coords[2] * S.dom[0], (long) cm.coords[1] * S.dom[1], (long) cm.coords[0] * S.dom[2]};
      const long sx = S.dom[0] + 2 * (GZ + PADDING), sy = S.dom[1] + 2 * (GZ + PADDING);
      for (long k = 0; k < S.dom[2]; ++k)
        for (long j = 0; j < S.dom[1]; ++j)
          for (long i = 0; i < S.dom[0]; ++i) {
            const double x = S.result[r][(i + GZ + PADDING) + (j + GZ + PADDING) * sx + (k + GZ + PADDING) * sx * sy];
            const double y = cur[o[0] + i + (o[1] + j) * G[0] + (o[2] + k) * G[0] * G[1]];
            const double rel = std::abs(x - y) / (std::abs(x) + std::abs(y) + 1e-300);
            worst = std::max(worst, rel);
            bad += !brick_detail::close_enough(x, y, BRICK_TOLERANCE);
          }
    }
    if (bad) {
      std::cout << "result mismatch!" << " (" << bad << " cells, worst relative difference " << worst << ")" << std::endl;
      return 2;
    }
    std::cout << "result match (worst relative difference " << worst << " after " << steps << " steps)" << std::endl;
  }
  return (failed || S.failures) ? 1 : 0;
}
