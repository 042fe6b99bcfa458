// drivers/strong.cpp -- strong-scaling driver: a fixed global domain cut into many small subdomains, Z-Morton ordered
// and dealt to the GPUs in contiguous id ranges (two-level decomposition).
//
// Same CLI letters and stdout lines as the reference's strong/main.cu + strong/args.cpp (-d/-s/-I/-v/-h; "Bri:",
// "calc :", "pack :", "call :", "wait :", "move :", "perf X GStencil/s", "part N"), same structure: ONE BrickDecomp of
// the subdomain size shared by all subdomains (strong/main.cu:131-134), Z-Morton neighbour lookup with periodic wrap
// (strong/args.cpp:36-56), section split (:104-113), one exchange + ST_ITER sweeps per period (:397-510).
// What differs: the reference copies same-rank ghosts with a cudaCopy kernel and packs/unpacks one MPI buffer per
// peer rank (strong/main.cu:188-395); here EVERY ghost region -- same GPU or not -- is one segment of a single pull
// plan (ghost <- owner's skin, read through NVLink peer pointers when the owner is another GPU): no pack, no unpack,
// one kernel.  All subdomains of a rank are swept by one launch (bk_stencil_apply_multi, strong/main.cu:85-99).
//
// Stitched mode (default whenever a rank's Z-Morton section is a box of subdomains -- always for power-of-two rank
// counts): the GPU analogue of the reference's mmap ghost aliasing (strong/main.cpp:205-262).  The rank's subdomains are
// presented to the marching kernel as ONE dense brick grid ("super grid") whose entries are global brick ids
// (subdomain * nbricks + local id) into the rank's two subdomain-major allocations: interior positions name the owning
// subdomain's interior brick, the one-brick shell names the ghost brick of the nearest boundary subdomain, and on an
// axis where the box spans the whole periodic domain the shell simply aliases the interior bricks of the far side.
// Same-GPU ghost regions are then never copied and never recomputed; only the regions on the box surface are pulled
// (from the owners' skins, over NVLink).  -M selects the per-subdomain launch instead.
//
// usage: strong [-d global_edge=512] [-s subdomain_edge=128] [-I periods=100] [-g gpus=1] [-S stencil] [-v]
#include <unistd.h>
#include <thread>
#include "common.h"

namespace {

struct Shared {
  int size = 1, iters = 100;
  unsigned dom_size = 512, sdom_size = 128;
  const StencilDef *st = nullptr;
  bool validate = false, no_stitch = false, no_overlap = false;
  unsigned long subdim = 4, allsubs = 64;
  std::vector<bElem *> base;  // rank -> device address of field 0 of its first subdomain
  std::vector<double> calc, call, wait, total, mbytes;
  std::vector<size_t> parts;
  std::vector<int> stitched;  // rank -> swept as one super grid?
  bElem *global_in = nullptr;
  std::vector<std::vector<bElem>> result;  // rank -> interiors of its subdomains, concatenated
};

// owner of a Z-Morton id and its index inside the owner (strong/args.cpp:104-113 + :47-55)
struct Sections {
  unsigned long len, shift, split;
  int size;
  Sections(unsigned long allsubs, int size) : size(size) {
    shift = allsubs % size;
    len = allsubs / size;
    split = shift * ((allsubs + size - 1) / size);
  }
  void range(int rank, unsigned long &l, unsigned long &r) const {
    const unsigned long mylen = len + (shift > (unsigned long) rank ? 1 : 0);
    l = (unsigned long) rank * mylen + (shift > (unsigned long) rank ? 0 : shift);
    r = l + mylen;
  }
  void owner(unsigned long id, int &dst, unsigned long &sub) const {
    if (id < split) dst = (int) (id / (len + 1)), sub = id % (len + 1);
    else dst = (int) ((id - shift) / len), sub = (id - shift) % len;
  }
};

unsigned long neighbour_id(unsigned long id, BitSet n, unsigned long subdim) {  // getrank's coordinate walk
  unsigned long c[3];
  bk_zmort_decode(id, c);
  for (int d = 0; d < 3; ++d) {
    long off = n.get(d + 1) ? 1 : n.get(-1 - d) ? -1 : 0;
    c[d] = (c[d] + subdim + off) % subdim;
  }
  return bk_zmort_encode(c);
}

void rank_main(int rank, Shared &S, Barrier &bar) {
  int ndev = 0;
  bkCheck(bk_device_count(&ndev));
  if (ndev == 0) throw std::runtime_error("no CUDA device: this build has no CPU path");
  const int dev = rank % ndev;
  bkCheck(bk_set_device(dev));
  const StencilDef *st = S.st;
  const long s = S.sdom_size, STRIDEG = s + 2 * GZ, STRIDEB = STRIDEG / TILE, STRIDE = STRIDEG + 2 * PADDING;
  const unsigned bSize = cal_size<BDIM>::value;

  BrickDecomp<3, BDIM> bDecomp({(unsigned) s, (unsigned) s, (unsigned) s}, GZ);
  bDecomp.initialize(skin3d_good);
  BrickInfo<3> bInfo = bDecomp.getBrickInfo();
  for (long k = 1; k < STRIDEB - 1; ++k)
    for (long j = 1; j < STRIDEB - 1; ++j)
      for (long i = 1; i < STRIDEB - 1; ++i) {
        const unsigned l = bDecomp[k][j][i];
        for (int id = 0; id < 27; ++id)
          if (bInfo.adj[bInfo.adj[l][id]][26 - id] != l) throw std::runtime_error("err");
      }
  const size_t nb = bInfo.nbricks, sub_elems = nb * bSize;

  Sections sec(S.allsubs, S.size);
  unsigned long mysec_l, mysec_r;
  sec.range(rank, mysec_l, mysec_r);
  const unsigned nsub = (unsigned) (mysec_r - mysec_l);

  // all subdomains of this rank live in two device allocations (field 0 / field 1), subdomain-major
  BrickInfo<3> bInfo_dev = movBrickInfo(bInfo, brickMemcpyHostToDevice);
  BrickStorage st0 = deviceBrickStorage((long) (nsub * nb), bSize), st1 = deviceBrickStorage((long) (nsub * nb), bSize);
  unsigned *grid_dev = nullptr;
  const std::vector<long> strideb = {STRIDEB, STRIDEB, STRIDEB};
  copyToDevice(strideb, grid_dev, bDecomp.gridData());

  if (S.validate) {  // inputs: interior cells of every subdomain, cut from the host's global array
    const std::vector<long> astride = {STRIDE, STRIDE, STRIDE};
    bElem *arr_dev = nullptr, *in_ptr = zeroArray(astride);
    copyToDevice(astride, arr_dev, in_ptr);
    const long G = S.dom_size;
    for (unsigned q = 0; q < nsub; ++q) {
      unsigned long c[3];
      bk_zmort_decode(mysec_l + q, c);
      for (long k = 0; k < s; ++k)
        for (long j = 0; j < s; ++j)
          std::memcpy(in_ptr + (PADDING + GZ) + (j + PADDING + GZ) * STRIDE + (k + PADDING + GZ) * STRIDE * STRIDE,
                      S.global_in + c[0] * s + (c[1] * s + j) * G + (c[2] * s + k) * G * G, s * sizeof(bElem));
      bkCheck(bk_memcpy_h2d(arr_dev, in_ptr, (size_t) STRIDE * STRIDE * STRIDE * sizeof(bElem), nullptr));
      BrickStorage view = st0;  // a Brick over subdomain q of field 0
      Brick3D b(&bInfo_dev, view, (unsigned) 0);
      b.dat = st0.dat.get() + q * sub_elems;
      copyToBrickDevice({STRIDEG, STRIDEG, STRIDEG}, {PADDING, PADDING, PADDING}, {0, 0, 0}, arr_dev, grid_dev, b);
      bkCheck(bk_device_sync());
    }
    bk_dev_free(arr_dev);
    free(in_ptr);
  } else {  // timing runs: the synthetic field is written straight into the bricks on the device (bk_fill_synthetic)
    const unsigned gd0[3] = {(unsigned) STRIDEB, (unsigned) STRIDEB, (unsigned) STRIDEB};
    const long glob[3] = {(long) S.dom_size, (long) S.dom_size, (long) S.dom_size};
    for (unsigned q = 0; q < nsub; ++q) {
      unsigned long c[3];
      bk_zmort_decode(mysec_l + q, c);
      const long org[3] = {(long) c[0] * s - GZ, (long) c[1] * s - GZ, (long) c[2] * s - GZ};
      bkCheck(bk_fill_synthetic(grid_dev, gd0, org, glob, 0x5EED, st0.dat.get() + q * sub_elems, bSize, nullptr));
    }
    bkCheck(bk_device_sync());
  }

  // per-subdomain field descriptors for the two sweep directions
  std::vector<bk_field_t> f01(nsub), f10(nsub);
  for (unsigned q = 0; q < nsub; ++q) {
    f01[q] = {&bInfo_dev.adj[0][0], st0.dat.get() + q * sub_elems, bSize, st1.dat.get() + q * sub_elems, bSize};
    f10[q] = {&bInfo_dev.adj[0][0], st1.dat.get() + q * sub_elems, bSize, st0.dat.get() + q * sub_elems, bSize};
  }
  bk_field_t *f01_dev = nullptr, *f10_dev = nullptr;
  copyToDevice({(long) nsub}, f01_dev, f01.data());
  copyToDevice({(long) nsub}, f10_dev, f10.data());

  S.base[rank] = st0.dat.get();
  bar.wait();

  // ---- stitched super grid (bk_stitch_*): is my section a box of subdomains? ---------------------------------------
  bk_stitch_box_t box;
  bkCheck(bk_stitch_box(mysec_l, nsub, S.subdim, &box));
  const bool stitched = !S.no_stitch && box.is_box && (unsigned long long) nsub * nb < (1ull << 32);
  const bool wrap[3] = {box.wrap[0] != 0, box.wrap[1] != 0, box.wrap[2] != 0};
  S.stitched[rank] = stitched;
  unsigned *sgrid_dev = nullptr;
  unsigned sdims[3];
  bkCheck(bk_stitch_dims(bDecomp.handle(), &box, sdims));
  std::vector<long> sgd = {(long) sdims[0], (long) sdims[1], (long) sdims[2]};
  if (stitched) {
    std::vector<unsigned> sg((size_t) sgd[0] * sgd[1] * sgd[2]);
    bkCheck(bk_stitch_grid(bDecomp.handle(), &box, sg.data()));
    copyToDevice(sgd, sgrid_dev, sg.data());
  }

  // link every ghost region to its owner's skin region (strong/main.cu:188-247), one pull plan for all of them.
  // Stitched: only regions on the surface of my box are ever read -- a region is needed iff every axis its direction
  // moves along leaves the box there (and the box does not wrap onto itself along it).
  std::vector<bk_seg_t> segs;
  size_t remote_bytes = 0, peers_mask = 0;
  for (unsigned q = 0; q < nsub; ++q)
    for (size_t i = 0; i < bDecomp.ghost.size(); ++i) {
      if (stitched && bk_stitch_region_needed(bDecomp.handle(), &box, mysec_l + q, (int) i) != 1) continue;
      int dst;
      unsigned long sub;
      sec.owner(neighbour_id(mysec_l + q, bDecomp.ghost[i].neighbor, S.subdim), dst, sub);
      bk_seg_t sg;
      sg.src = S.base[dst] + sub * sub_elems + (size_t) bDecomp.skin[i].pos * bSize;
      sg.dst = st0.dat.get() + q * sub_elems + (size_t) bDecomp.ghost[i].pos * bSize;
      sg.bytes = (size_t) bDecomp.ghost[i].len * bSize * sizeof(bElem);
      segs.push_back(sg);
      if (dst != rank) {
        remote_bytes += sg.bytes, peers_mask |= 1ull << dst;
        if (dst % ndev != dev) bkCheck(bk_peer_enable(dst % ndev));  // no P2P between the two GPUs: fail here, not in the pull kernel
      }
    }
  ExchangeView ev(segs);

  void *evDone, *c0, *c1, *x0, *x1;
  for (void **e : {&evDone, &c0, &c1, &x0, &x1}) bkCheck(bk_event_create(e));
  const unsigned gd[3] = {(unsigned) STRIDEB, (unsigned) STRIDEB, (unsigned) STRIDEB};
  const unsigned full_lo[3] = {0, 0, 0}, full_hi[3] = {gd[0], gd[1], gd[2]};
  const unsigned skip_lo[3] = {1, 1, 1}, skip_hi[3] = {gd[0] - 1, gd[1] - 1, gd[2] - 1};
  double calctime = 0, calltime = 0, waittime = 0;
  // stitched sweeps: whole-allocation views, sweep box per axis, passes of `fuse` time steps
  BrickStorage vA = st0, vB = st1;
  Brick3D sA(&bInfo_dev, vA, (unsigned) 0), sB(&bInfo_dev, vB, (unsigned) 0);
  std::vector<long> sw_lo(3), sw_hi(3), own_lo(3), own_hi(3);
  for (int d = 0; d < 3; ++d) {
    own_lo[d] = 1, own_hi[d] = sgd[d] - 1;
    sw_lo[d] = wrap[d] ? 1 : 0, sw_hi[d] = wrap[d] ? sgd[d] - 1 : sgd[d];
  }
  int fuse = bk_stencil_fused_steps(st->id);
  if (fuse < 1 || st->st_iter % fuse || (st->st_iter / fuse) % 2) fuse = 1;
  const int npass = st->st_iter / fuse;

  // the exchange runs on its own high-priority stream so that the first pass can overlap it (stitched mode): the CTAs of
  // pass 0 that read only bricks I own start at once on the compute stream (BK_PART_READY), the CTAs that touch the
  // box surface's ghost bricks follow the pull on the exchange stream (BK_PART_REST) -- the weak driver's scheme
  void *comm_stream, *evX;
  bkCheck(bk_stream_create_priority(&comm_stream, 1));
  bkCheck(bk_event_create(&evX));
  std::vector<long> rdy_lo(3), rdy_hi(3);  // bricks that are final without the exchange: everything but real ghost shell
  for (int d = 0; d < 3; ++d) rdy_lo[d] = wrap[d] ? 0 : 1, rdy_hi[d] = wrap[d] ? sgd[d] : sgd[d] - 1;
  bool overlap = stitched && !S.no_overlap;
  auto stitched_pass = [&](int p, int part, void *stream) -> bool {
    const bool last = p == npass - 1;
    Brick3D &src = (p % 2) ? sB : sA, &dst = (p % 2) ? sA : sB;
    // the stitched grid aliases ghost positions onto other subdomains' bricks on purpose: it, not the per-subdomain
    // adjacency list, is the topology of this launch (BK_PART_GRID_TOPOLOGY)
    return brickAdvance(st->id, fuse, sgrid_dev, sgd, src, dst, last ? own_lo : sw_lo, last ? own_hi : sw_hi, rdy_lo, rdy_hi,
                        part | BK_PART_GRID_TOPOLOGY, nullptr, stream);
  };

  auto brick_func = [&]() {
    bkCheck(bk_event_record(evDone, nullptr));
    bkCheck(bk_event_sync(evDone));
    double t0 = omp_get_wtime();
    bar.wait();  // every rank's skins are final
    waittime += omp_get_wtime() - t0;
    float ms = 0;
    if (overlap && npass > 1) {
      bkCheck(bk_event_record(x0, comm_stream));
      ev.exchange(comm_stream);
      bkCheck(bk_event_record(x1, comm_stream));
      bkCheck(bk_event_record(c0, nullptr));
      if (stitched_pass(0, BK_PART_READY, nullptr)) {
        if (!stitched_pass(0, BK_PART_REST, comm_stream)) throw std::runtime_error("split pass unavailable");
        bkCheck(bk_event_record(evX, comm_stream));
        bkCheck(bk_stream_wait_event(nullptr, evX));
      } else {  // no split kernel for this layout: the whole pass after the pull
        overlap = false;
        bkCheck(bk_event_record(evX, comm_stream));
        bkCheck(bk_stream_wait_event(nullptr, evX));
        if (!stitched_pass(0, BK_PART_ALL, nullptr)) throw std::runtime_error("marching kernel unavailable for the stitched grid: rerun with -M");
      }
      // pass 1 overwrites the storage whose skins the neighbours are pulling: wait until every pull has finished
      bkCheck(bk_event_sync(evX));
      bkCheck(bk_event_elapsed_ms(x0, x1, &ms));
      calltime += ms / 1e3;
      t0 = omp_get_wtime();
      bar.wait();
      waittime += omp_get_wtime() - t0;
      for (int p = 1; p < npass; ++p)
        if (!stitched_pass(p, BK_PART_ALL, nullptr)) throw std::runtime_error("pass unavailable");
      bkCheck(bk_event_record(c1, nullptr));
      bkCheck(bk_event_sync(c1));
      bkCheck(bk_event_elapsed_ms(c0, c1, &ms));
      calctime += ms / 1e3;  // includes the overlapped pull
      return;
    }
    bkCheck(bk_event_record(x0, nullptr));
    ev.exchange(nullptr);
    bkCheck(bk_event_record(x1, nullptr));
    bkCheck(bk_event_sync(x1));
    bkCheck(bk_event_elapsed_ms(x0, x1, &ms));
    calltime += ms / 1e3;
    t0 = omp_get_wtime();
    bar.wait();  // every pull has finished before a skin is overwritten
    waittime += omp_get_wtime() - t0;
    bkCheck(bk_event_record(c0, nullptr));
    if (stitched) {
      // one sweep (or fused pass of two time steps) over the whole super grid; the shell is swept only along axes whose
      // shell is real ghost storage (communication avoiding, weak/main.cu:275-285), never where it aliases my interior
      for (int p = 0; p < npass; ++p)
        if (!stitched_pass(p, BK_PART_ALL, nullptr))
          throw std::runtime_error("marching kernel unavailable for the stitched grid: rerun with -M");
    } else {
      for (int sw = 0; sw < st->st_iter; ++sw) {
        const bool last = sw == st->st_iter - 1;
        bkCheck(bk_stencil_apply_multi(st->id, (sw % 2) ? f10_dev : f01_dev, nsub, grid_dev, gd, last ? skip_lo : full_lo,
                                       last ? skip_hi : full_hi, nullptr, nullptr));
      }
    }
    bkCheck(bk_event_record(c1, nullptr));
    bkCheck(bk_event_sync(c1));
    bkCheck(bk_event_elapsed_ms(c0, c1, &ms));
    calctime += ms / 1e3;
  };

  brick_func();
  calctime = calltime = waittime = 0;
  bar.wait();
  const double t_start = omp_get_wtime();
  for (int i = 0; i < S.iters; ++i) brick_func();
  bkCheck(bk_device_sync());
  bar.wait();
  const double per_period = (omp_get_wtime() - t_start) / S.iters;
  const int cnt = S.iters * st->st_iter;
  S.calc[rank] = calctime / cnt, S.call[rank] = calltime / cnt, S.wait[rank] = waittime / cnt;
  S.total[rank] = per_period / st->st_iter;
  S.mbytes[rank] = remote_bytes * 2 * 1e-6;
  S.parts[rank] = 2 * (size_t) __builtin_popcountll(peers_mask);

  if (S.validate) {
    const std::vector<long> astride = {STRIDE, STRIDE, STRIDE};
    bElem *arr_dev = nullptr, *host = zeroArray(astride);
    copyToDevice(astride, arr_dev, host);
    S.result[rank].resize((size_t) nsub * s * s * s);
    for (unsigned q = 0; q < nsub; ++q) {
      BrickStorage view = st0;
      Brick3D b(&bInfo_dev, view, (unsigned) 0);
      b.dat = (((stitched ? npass : st->st_iter) % 2) ? st1 : st0).dat.get() + q * sub_elems;
      copyFromBrickDevice({s, s, s}, {PADDING, PADDING, PADDING}, {GZ, GZ, GZ}, arr_dev, grid_dev, b);
      copyFromDevice(astride, host, arr_dev);
      for (long k = 0; k < s; ++k)
        for (long j = 0; j < s; ++j)
          std::memcpy(&S.result[rank][(size_t) q * s * s * s + (k * s + j) * s],
                      host + (PADDING + GZ) + (j + PADDING + GZ) * STRIDE + (k + PADDING + GZ) * STRIDE * STRIDE, s * sizeof(bElem));
    }
    bk_dev_free(arr_dev);
    free(host);
  }
  bar.wait();
  if (rank == 0) {
    mpi_stats calc_s = mpi_statistics(S.calc), call_s = mpi_statistics(S.call), wait_s = mpi_statistics(S.wait);
    mpi_stats tot_s = mpi_statistics(S.total), size_s = mpi_statistics(S.mbytes);
    std::vector<double> zero(S.size, 0.0), spd(S.size);
    for (int r = 0; r < S.size; ++r) spd[r] = S.mbytes[r] * 1e-3 / std::max(1e-12, (S.call[r] + S.wait[r]) * st->st_iter);
    const double total = tot_s.max;
    std::cout << "Bri: " << total << " : " << calc_s.avg + call_s.avg + wait_s.avg << std::endl;
    std::cout << "calc : " << calc_s << std::endl;
    std::cout << "pack : " << mpi_statistics(zero) << std::endl;
    std::cout << "  | pack speed (GB/s): " << mpi_statistics(zero) << std::endl;
    std::cout << "call : " << call_s << std::endl;
    std::cout << "wait : " << wait_s << std::endl;
    std::cout << "  | MPI size (MB): " << size_s << std::endl;
    std::cout << "  | MPI speed (GB/s): " << mpi_statistics(spd) << std::endl;
    std::cout << "move : " << mpi_statistics(zero) << std::endl;
    double perf = S.dom_size / 1000.0;
    perf = perf * perf * perf / total;
    std::cout << "perf " << perf << " GStencil/s" << std::endl;
    size_t parts = 0;
    for (size_t p : S.parts) parts = std::max(parts, p);
    std::cout << "part " << parts << std::endl;
    int ns = 0;
    for (int v : S.stitched) ns += v;
    std::cout << "stitched ranks " << ns << " of " << S.size << std::endl;
  }
  freeBrickInfoDevice(bInfo_dev);
  bk_dev_free(grid_dev);
  if (sgrid_dev) bk_dev_free(sgrid_dev);
  bk_dev_free(f01_dev);
  bk_dev_free(f10_dev);
}

}  // namespace

int main(int argc, char **argv) {
  Shared S;
  std::string sname = "mpi7pt";
  int c;
  if (const char *e = getenv("BRICK_RANKS")) S.size = atoi(e);
  while ((c = getopt(argc, argv, "d:s:I:g:S:vhMO")) != -1) switch (c) {
      case 'd': S.dom_size = std::stoi(optarg); break;
      case 's': S.sdom_size = std::stoi(optarg); break;
      case 'I': S.iters = std::stoi(optarg); break;
      case 'g': S.size = std::stoi(optarg); break;
      case 'S': sname = optarg; break;
      case 'v': S.validate = true; break;
      case 'M': S.no_stitch = true; break;
      case 'O': S.no_overlap = true; break;
      default:
        printf("Program options\n  -h: help\n  -O: do not overlap the exchange with the first pass (stitched mode)\n  -M: per-subdomain launches and ghost copies instead of the stitched super grid\n  -d n: global domain edge (default 512)\n  -s n: subdomain edge (default 128)\n"
               "  -I n: exchange periods (default 100)\n  -g n: GPUs = ranks (default 1)\n  -S name: mpi7pt mpi13pt mpi25pt mpi125pt\n"
               "  -v: validate against a CPU sweep of the global periodic array\n");
        return 0;
    }
  S.st = find_stencil(sname);
  S.subdim = S.dom_size / S.sdom_size;
  if (!S.st || S.st->id == BK_ST_7PT || S.subdim * S.sdom_size != S.dom_size || (S.subdim & (S.subdim - 1)) || S.sdom_size % TILE ||
      S.sdom_size < 2 * GZ) {
    std::cerr << "need an MPI stencil, d/s a power of two, s a multiple of " << TILE << " and >= " << 2 * GZ << std::endl;
    return 1;
  }
  S.allsubs = S.subdim * S.subdim * S.subdim;
  std::cout << "Pagesize " << sysconf(_SC_PAGESIZE) << "; MPI Size " << S.size << " * OpenMP threads " << omp_get_max_threads() << std::endl;
  std::cout << "Domain size of " << S.dom_size << "^3 decomposed into " << S.sdom_size << "^3 subdomains" << std::endl;
  std::cout << "Total of " << S.allsubs << " subdomains, " << (S.allsubs + S.size - 1) / S.size << " per rank" << std::endl;

  S.base.assign(S.size, nullptr);
  S.calc.assign(S.size, 0), S.call = S.wait = S.total = S.mbytes = S.calc;
  S.parts.assign(S.size, 0);
  S.stitched.assign(S.size, 0);
  S.result.resize(S.size);
  const long G[3] = {(long) S.dom_size, (long) S.dom_size, (long) S.dom_size};
  std::vector<bElem> initial;
  if (S.validate) {
    S.global_in = randomArray({G[0], G[1], G[2]});
    initial.assign(S.global_in, S.global_in + (size_t) G[0] * G[1] * G[2]);
  }
  Barrier bar(S.size);
  std::vector<std::thread> th;
  for (int r = 0; r < S.size; ++r)
    th.emplace_back([&, r] {
      try {
        rank_main(r, S, bar);
      } catch (const std::exception &e) {
        std::cerr << "rank " << r << ": " << e.what() << std::endl;
        exit(EXIT_FAILURE);
      }
    });
  for (auto &t : th) t.join();

  if (S.validate) {
    const int steps = (S.iters + 1) * S.st->st_iter;
    cpu_periodic_steps(S.global_in, G, stencil_taps(S.st->id, nullptr), S.st->radius, steps);
    Sections sec(S.allsubs, S.size);
    long bad = 0;
    double worst = 0;
    const long s = S.sdom_size;
    for (int r = 0; r < S.size; ++r) {
      unsigned long l, rr;
      sec.range(r, l, rr);
      for (unsigned long id = l; id < rr; ++id) {
        unsigned long co[3];
        bk_zmort_decode(id, co);
        for (long k = 0; k < s; ++k)
          for (long j = 0; j < s; ++j)
            for (long i = 0; i < s; ++i) {
              const double x = S.result[r][(size_t) (id - l) * s * s * s + (k * s + j) * s + i];
              const double y = S.global_in[co[0] * s + i + (co[1] * s + j) * G[0] + (co[2] * s + k) * G[0] * G[1]];
              worst = std::max(worst, std::abs(x - y) / (std::abs(x) + std::abs(y) + 1e-300));
              bad += !brick_detail::close_enough(x, y, BRICK_TOLERANCE);
            }
      }
    }
    if (bad) {
      std::cout << "result mismatch! (" << bad << " cells, worst relative difference " << worst << ")" << std::endl;
      return 2;
    }
    std::cout << "result match (worst relative difference " << worst << " after " << steps << " steps)" << std::endl;
  }
  return 0;
}
