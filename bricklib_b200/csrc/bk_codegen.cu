// bk_codegen.cu -- the stencil-script back end: CUDA source GENERATED per stencil, compiled at run time with NVRTC.
//
// The reference turns every stencils/*.py expression into specialised vector code at build time (codegen/vecscatter
// :145-175 drives codegen/st/codegen/backend/cuda.py:123-154: one warp per brick, shuffles between lanes).  This file
// plays that role for sm_100a: from a lowered tap list  out(i,j,k) = post(sum_t c_t * pre(in(i+di_t, j+dj_t, k+dk_t)))
// it EMITS the text of a marching kernel specialised to the taps -- same structure as the hand-written kernels of
// bk_stencil_tiled.cu (a CTA owns a tile of bricks and marches along k; producer warps fill a shared-memory ring with
// cp.async.bulk; consumer threads own an x-pair x YT rows patch and keep one partial output per live k plane in
// registers) -- with the tap loop fully unrolled: every shared-memory row is loaded once per plane and feeds straight
// line FMAs whose row / column / k-slot are literals.  Only the coefficient VALUES stay run-time data (a by-value kernel
// parameter, read from the constant bank), so one compiled kernel serves every coefficient set of a tap pattern.
// NVRTC (dlopen'ed: the library loads without it) compiles the text for sm_100a into a cubin, the runtime's library API
// loads it (cudaLibraryLoadData / cudaLibraryGetKernel), launches go through the same CTA enumeration as the
// compiled-in kernels (bk::launch_generated), so split launches for exchange overlap work unchanged.
//
// Anything the generator declines (radius > 4, more taps than fit the parameter space, NVRTC missing) stays on the
// per-brick tap-table kernel k_taps (bk_stencil.cu).
#include "bk_common.h"
#include "bk_codegen.h"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace {

const char *kCommonHeader =
#include "bk_march_common.inc"
    ;

// ---- NVRTC through dlopen ---------------------------------------------------------------------------------------------
struct Nvrtc {
  void *h = nullptr;
  int (*createProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*compileProgram)(void *, int, const char *const *) = nullptr;
  int (*getCUBINSize)(void *, size_t *) = nullptr;
  int (*getCUBIN)(void *, char *) = nullptr;
  int (*getProgramLogSize)(void *, size_t *) = nullptr;
  int (*getProgramLog)(void *, char *) = nullptr;
  int (*destroyProgram)(void **) = nullptr;
  bool ok = false;
};

Nvrtc &nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      n.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.h) break;
    }
    if (!n.h) return;
    auto sym = [&](const char *s) { return dlsym(n.h, s); };
    n.createProgram = (decltype(n.createProgram)) sym("nvrtcCreateProgram");
    n.compileProgram = (decltype(n.compileProgram)) sym("nvrtcCompileProgram");
    n.getCUBINSize = (decltype(n.getCUBINSize)) sym("nvrtcGetCUBINSize");
    n.getCUBIN = (decltype(n.getCUBIN)) sym("nvrtcGetCUBIN");
    n.getProgramLogSize = (decltype(n.getProgramLogSize)) sym("nvrtcGetProgramLogSize");
    n.getProgramLog = (decltype(n.getProgramLog)) sym("nvrtcGetProgramLog");
    n.destroyProgram = (decltype(n.destroyProgram)) sym("nvrtcDestroyProgram");
    n.ok = n.createProgram && n.compileProgram && n.getCUBINSize && n.getCUBIN && n.getProgramLogSize && n.getProgramLog &&
           n.destroyProgram;
  });
  return n;
}

int floordiv2(int e) { return e >= 0 ? e / 2 : -((1 - e) / 2); }

std::string lit(double v) {
  char b[64];
  snprintf(b, sizeof(b), "%.17g", v);
  std::string s(b);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}

std::string pw_expr(const std::string &x, const bk_pointwise_t &p) {
  switch (p.op) {
    case BK_OP_MAX: return "fmax(" + x + ", " + lit(p.c) + ")";
    case BK_OP_MIN: return "fmin(" + x + ", " + lit(p.c) + ")";
    case BK_OP_ABS: return "fabs(" + x + ")";
  }
  return x;
}

}  // namespace

namespace bk {

struct GenStencil {
  std::vector<GenTap> taps;
  bk_pointwise_t pre, post;
  int RX = 0, RY = 0, ZLO = 0, ZHI = 0;  // extents of the taps: |di| <= RX, |dj| <= RY, ZLO <= dk <= ZHI
  bool corners = false;                  // some tap moves along i AND j: the j-halo rows need the corner bricks too
  GenGeom geom;
  std::string source, log;
  std::vector<char> cubin;
  std::vector<double> coef;              // kernel parameter #2, by value
  std::mutex mu;
  std::map<int, std::pair<cudaLibrary_t, cudaKernel_t>> loaded;  // per device
};

namespace {

// fixed geometry of the generated kernels: 4x4-brick tiles, 2 rows per consumer thread (8 consumer warps), 2 producer
// warps, two planes per ring stage, 3 stages
constexpr int TI = 4, TJ = 4, YT = 2, G = 2, D = 3, NPW = 2;
constexpr int SW = TI + 2, SH = TJ + 1, SLOTP = G * 512 + 64;
constexpr int STAGE = ((SH * SW * SLOTP + 127) / 128) * 128;
constexpr int NCONS = TI * TJ * 32 / YT, NCW = NCONS / 32, NT = NCONS + 32 * NPW;

void emit(GenStencil &g) {
  const int R = std::max({g.RX, g.RY, -g.ZLO, g.ZHI, 1});
  const int RUP = ((R + G - 1) / G) * G;
  const int W = g.ZHI - g.ZLO + 1;
  const int CX = (g.RX + 1) / 2;  // 16-byte chunks of i-halo per side
  const int NJH = g.corners ? TI + 2 : TI;
  const int ntap = (int) g.taps.size();
  // CTAs per SM the register allocation must allow: two for small tap sets (<= 102 registers: 16 consumer warps per SM hide
  // the shared-memory latency), one for large ones (up to 168 registers for the unrolled FMA block)
  int minb = ntap <= 40 ? 2 : 1;
  if (const char *e = getenv("BK_GEN_MINB")) minb = atoi(e) == 2 ? 2 : 1;  // developer knob
  std::ostringstream o;
  o << "// generated by libbrick_b200 (bk_codegen.cu): marching kernel specialised to " << ntap << " taps, |di| <= " << g.RX
    << ", |dj| <= " << g.RY << ", " << g.ZLO << " <= dk <= " << g.ZHI << "\n";
  o << kCommonHeader << "\n";
  o << "struct GenCoef { double c[" << ntap << "]; };\n";
  o << "#define TI " << TI << "\n#define TJ " << TJ << "\n#define YT " << YT << "\n#define GP " << G << "\n#define DS " << D
    << "\n#define NPW " << NPW << "\n#define SW " << SW << "\n#define SLOTP " << SLOTP << "\n#define STAGE " << STAGE
    << "\n#define NCONS " << NCONS << "\n#define NCW " << NCW << "\n#define RUP " << RUP << "\n#define RY " << g.RY
    << "\n#define MINB " << minb << "\n#define NJH " << NJH << "\n#define CORNER0 " << (g.corners ? 0 : 1) << "\n#define WS " << W << "\n#define ZHI " << g.ZHI
    << "\n";
  o << R"SRC(
__device__ __forceinline__ int slotoff(int bi, int bj) { return (bj * SW + bi) * SLOTP; }

extern "C" __global__ void __launch_bounds__(NCONS + 32 * NPW, MINB) bk_gen(const __grid_constant__ TiledArgs a,
                                                                      const __grid_constant__ GenCoef cf) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  const unsigned ring_u32 = smem_u32(ring);
  const unsigned bar_full = ring_u32 + DS * STAGE;
  const unsigned bar_empty = bar_full + DS * 8;
  const double *fin = a.in;
  double *fout = a.out;
  const unsigned long long in_step = a.in_step, out_step = a.out_step;
  const int tid = threadIdx.x;
  int bq = 0, brel = (int) blockIdx.x;
  while (bq + 1 < a.nbox && brel >= a.box[bq + 1].first) ++bq;
  brel -= a.box[bq].first;
  const int tx = a.box[bq].lo[0] + brel % a.box[bq].dim[0];
  brel /= a.box[bq].dim[0];
  const int ty = a.box[bq].lo[1] + brel % a.box[bq].dim[1];
  const int tseg = a.box[bq].lo[2] + brel / a.box[bq].dim[1];
  const int i0 = a.lo[0] + tx * TI, j0 = a.lo[1] + ty * TJ;
  int kb0, nl;
  seg_range(a, tseg, kb0, nl);
  const int P = nl * 8 + 2 * RUP;  // planes streamed; plane n is absolute plane kb0*8 - RUP + n
  const int NS = P / GP;

  if (tid == 0) {
    for (int s = 0; s < DS; ++s) {
      mbar_init(bar_full + 8 * s, 32 * NPW);
      mbar_init(bar_empty + 8 * s, NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= NCONS) {
    // ---- producer warps: one copy job per lane (own bricks, i-halo bricks, j-halo rows), as in bk_stencil_tiled.cu ----
    const int job0 = ((tid - NCONS) >> 5) + NPW * (tid & 31);
    int job = job0, kind = 0, sbi = 0, sbj = 0;
    if (job < TI * TJ) {
      kind = 1, sbi = 1 + job % TI, sbj = 1 + job / TI;
    } else if ((job -= TI * TJ) < 2 * TJ) {
      kind = 1, sbi = (job & 1) ? TI + 1 : 0, sbj = 1 + (job >> 1);
    } else if ((job -= 2 * TJ) < 2 * NJH) {
      kind = 2 + (job & 1), sbi = CORNER0 + (job >> 1), sbj = (job & 1) ? TJ + 1 : 0;
    }
    if (RY == 0 && kind >= 2) kind = 0;
    const unsigned dsto = slotoff(sbi, kind >= 2 ? 0 : sbj) + (kind == 2 ? (8 - RY) * 64 : 0);
    const int gi = i0 + sbi - 1, gj = j0 + sbj - 1;
    const bool inside = kind != 0 && gi >= 0 && gi < a.gx && gj >= 0 && gj < a.gy;
    const int z_first = kb0 * 8 - RUP;
    int kb = (z_first >= 0) ? z_first / 8 : -((7 - z_first) / 8);
    int pz = z_first - kb * 8;
    auto brick_id = [&](int k) -> unsigned {
      return (inside && k >= 0 && k < a.gz) ? __ldg(a.grid + ((size_t) k * a.gy + gj) * a.gx + gi) : 0u;
    };
    unsigned idn = brick_id(kb), idc = 0;
    const unsigned bytes = kind == 0 ? 0 : kind == 1 ? GP * 512 : GP * RY * 64;
    int st = 0;
    unsigned ph = 0;
    bool fresh = true;
    for (int n = 0; n < NS; ++n) {
      if (fresh) idc = idn, idn = brick_id(kb + 1), fresh = false;
      if (n >= DS) mbar_wait(bar_empty + 8 * st, ph ^ 1);
      const unsigned fb = bar_full + 8 * st;
      mbar_expect_tx(fb, bytes);
      const unsigned sb = ring_u32 + st * STAGE;
      const double *src = fin + (size_t) idc * in_step + pz * 64;
      if (kind == 1) {
        bulk_g2s(sb + dsto, src, GP * 512, fb);
      } else if (kind >= 2) {
        src += (kind == 2 ? (8 - RY) * 8 : 0);
#pragma unroll
        for (int g = 0; g < GP; ++g) bulk_g2s(sb + dsto + g * 512, src + g * 64, RY * 64, fb);
      }
      pz += GP;
      if (pz == 8) pz = 0, ++kb, fresh = true;
      if (++st == DS) st = 0, ph ^= 1;
    }
    return;
  }

  // ---- consumers: x-pair (cells 2c, 2c+1) x YT rows of one brick ----------------------------------------------------------
  const int c = tid & 3, e = (tid >> 2) & 1;
  int rest = tid >> 3;
  const int y0 = (rest % (8 / YT)) * YT;
  rest /= (8 / YT);
  const int bi = (rest % (TI / 2)) * 2 + e, bj = rest / (TI / 2);
  const int own_slot = slotoff(bi + 1, bj + 1);
  const int own_off = own_slot + y0 * 64 + c * 16;
  int joff[2 * RY + 1];  // rows y0-RY .. y0-1, then y0+YT .. y0+YT+RY-1 (byte offsets from the plane base)
#pragma unroll
  for (int h = 0; h < 2 * RY; ++h) {
    const int ya = (h < RY) ? y0 - RY + h : y0 + YT + (h - RY);
    int base;
    if (ya < 0) base = slotoff(bi + 1, bj) + (8 + ya) * 64;
    else if (ya >= 8) base = slotoff(bi + 1, (bj == TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
    else base = own_slot + ya * 64;
    joff[h] = base + c * 16;
  }
  // chunk m of a row (cells 2(c+m), 2(c+m)+1): inside my brick, or in the slot of the brick to the left / right
  auto ioff = [&](int m) -> int { return 16 * m + ((c + m < 0) ? 64 - SLOTP : (c + m > 3) ? SLOTP - 64 : 0); };
  const bool mine = (i0 + bi < a.hi[0]) && (j0 + bj < a.hi[1]);
  const unsigned *gcol = a.grid + ((size_t) kb0 * a.gy + (j0 + bj)) * a.gx + (i0 + bi);
  const size_t glayer = (size_t) a.gy * a.gx;
  unsigned id_next = mine ? __ldg(gcol) : 0u;
  double *outp = fout;
  double2 acc[WS][YT];
#pragma unroll
  for (int w = 0; w < WS; ++w)
#pragma unroll
    for (int r = 0; r < YT; ++r) acc[w][r] = make_double2(0.0, 0.0);
  int st = 0, pl = 0;
  unsigned ph = 0;
  int orel = -RUP - ZHI;  // output plane completed by the current input plane, relative to the segment start
  const int nout = nl * 8;
)SRC";
  // per-thread chunk offsets that the taps use
  for (int m = -CX; m <= CX; ++m)
    if (m) o << "  const int io" << (m < 0 ? "m" : "p") << std::abs(m) << " = ioff(" << m << ");\n";
  o << "#pragma unroll 1\n  for (int tb = 0; tb < P; tb += WS) {\n";
  for (int u = 0; u < W; ++u) {
    auto slot = [&](int dz) { return (((u - dz) % W) + W) % W; };
    const int sF = slot(g.ZHI), sN = slot(g.ZLO);
    o << "    if (tb + " << u << " < P) {\n";
    o << "      if (pl == 0) mbar_wait(bar_full + 8 * st, ph);\n";
    o << "      const unsigned char *pb = ring + st * STAGE + pl * 512;\n";
    for (int r = 0; r < YT; ++r) o << "      acc[" << sN << "][" << r << "] = make_double2(0.0, 0.0);\n";
    for (int rr = 0; rr < YT + 2 * g.RY; ++rr) {
      const int yrel = rr - g.RY;  // row y0 + yrel
      // the taps that read this row for some own row r: r + dj == yrel
      std::vector<std::pair<int, int>> use;  // (tap index, r)
      for (int t = 0; t < ntap; ++t) {
        const int r = yrel - g.taps[t].dj;
        if (r >= 0 && r < YT) use.push_back({t, r});
      }
      if (use.empty()) continue;
      std::vector<bool> need(2 * CX + 1, false);
      for (auto &tu : use)
        for (int cell = 0; cell < 2; ++cell) need[floordiv2(g.taps[tu.first].di + cell) + CX] = true;
      o << "      {\n        const unsigned char *pr = pb + ";
      if (yrel < 0) o << "joff[" << rr << "]";
      else if (yrel >= YT) o << "joff[" << (rr - YT) << "]";
      else o << "own_off + " << yrel * 64;
      o << ";\n";
      for (int m = -CX; m <= CX; ++m) {
        if (!need[m + CX]) continue;
        const std::string nm = std::string("q") + (m < 0 ? "m" : m > 0 ? "p" : "c") + (m ? std::to_string(std::abs(m)) : "");
        o << "        double2 " << nm << " = *reinterpret_cast<const double2 *>(pr" << (m ? std::string(" + io") + (m < 0 ? "m" : "p") + std::to_string(std::abs(m)) : "") << ");\n";
        if (g.pre.op != BK_OP_NONE)
          o << "        " << nm << ".x = " << pw_expr(nm + ".x", g.pre) << ", " << nm << ".y = " << pw_expr(nm + ".y", g.pre) << ";\n";
      }
      for (auto &tu : use) {
        const GenTap &tp = g.taps[tu.first];
        const int s = slot(tp.dk);
        for (int cell = 0; cell < 2; ++cell) {
          const int ecell = tp.di + cell, m = floordiv2(ecell), comp = ecell - 2 * m;
          const std::string nm = std::string("q") + (m < 0 ? "m" : m > 0 ? "p" : "c") + (m ? std::to_string(std::abs(m)) : "");
          const std::string dst = "acc[" + std::to_string(s) + "][" + std::to_string(tu.second) + "]." + (cell ? "y" : "x");
          o << "        " << dst << " = fma(cf.c[" << tu.first << "], " << nm << "." << (comp ? "y" : "x") << ", " << dst << ");\n";
        }
      }
      o << "      }\n";
    }
    // output plane orel is complete
    o << "      if (orel >= 0 && orel < nout) {\n        const int oz = orel & 7;\n        if (oz == 0) {\n"
         "          outp = fout + (size_t) id_next * out_step + y0 * 8 + c * 2;\n"
         "          if (mine && (orel >> 3) + 1 < nl) id_next = __ldg(gcol + ((orel >> 3) + 1) * glayer);\n        }\n"
         "        if (mine) {\n";
    for (int r = 0; r < YT; ++r) {
      const std::string v = "acc[" + std::to_string(sF) + "][" + std::to_string(r) + "]";
      o << "          *reinterpret_cast<double2 *>(outp + oz * 64 + " << r * 8 << ") = make_double2(" << pw_expr(v + ".x", g.post) << ", "
        << pw_expr(v + ".y", g.post) << ");\n";
    }
    o << "        }\n      }\n";
    o << "      ++orel;\n      if (++pl == GP) {\n        pl = 0;\n        __syncwarp();\n        if ((tid & 31) == 0) mbar_arrive(bar_empty + 8 * st);\n"
         "        if (++st == DS) st = 0, ph ^= 1;\n      }\n    }\n";
  }
  o << "  }\n}\n";
  g.source = o.str();
  g.geom.TI = TI, g.geom.TJ = TJ, g.geom.ovh = 2 * RUP + 2, g.geom.threads = NT;
  g.geom.smem = (size_t) D * STAGE + 2 * D * 8 + 128;
}

bool compile(GenStencil &g) {
  Nvrtc &n = nvrtc();
  if (!n.ok) {
    g.log = "libnvrtc is not available";
    return false;
  }
  void *prog = nullptr;
  if (n.createProgram(&prog, g.source.c_str(), "bk_gen.cu", 0, nullptr, nullptr) != 0) {
    g.log = "nvrtcCreateProgram failed";
    return false;
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
  const int rc = n.compileProgram(prog, 4, opts);
  size_t ls = 0;
  n.getProgramLogSize(prog, &ls);
  if (ls > 1) {
    g.log.resize(ls);
    n.getProgramLog(prog, &g.log[0]);
  }
  bool ok = rc == 0;
  if (ok) {
    size_t cs = 0;
    ok = n.getCUBINSize(prog, &cs) == 0 && cs > 0;
    if (ok) {
      g.cubin.resize(cs);
      ok = n.getCUBIN(prog, g.cubin.data()) == 0;
    }
  }
  n.destroyProgram(&prog);
  return ok;
}

}  // namespace

GenStencil *gen_create(const std::vector<GenTap> &taps, const bk_pointwise_t &pre, const bk_pointwise_t &post, std::string *why) {
  GenStencil *g = new GenStencil();
  g->taps = taps, g->pre = pre, g->post = post;
  g->ZLO = 1 << 20, g->ZHI = -(1 << 20);
  for (const GenTap &t : taps) {
    g->RX = std::max(g->RX, std::abs(t.di)), g->RY = std::max(g->RY, std::abs(t.dj));
    g->ZLO = std::min(g->ZLO, t.dk), g->ZHI = std::max(g->ZHI, t.dk);
    g->corners = g->corners || (t.di != 0 && t.dj != 0);
    g->coef.push_back(t.c);
  }
  auto fail = [&](const std::string &m) {
    if (why) *why = m;
    delete g;
    return (GenStencil *) nullptr;
  };
  if (taps.empty() || g->RX > 4 || g->RY > 4 || g->ZLO < -4 || g->ZHI > 4) return fail("radius beyond half a brick");
  if (taps.size() * sizeof(double) + tiled_args_bytes() > 3800) return fail("too many taps for the kernel parameter space");
  emit(*g);
  if (!compile(*g)) return fail("NVRTC: " + g->log);
  if (const char *dir = getenv("BK_GEN_DUMP")) {  // developer knob: keep the generated text and the cubin for cuobjdump
    static std::atomic<int> serial{0};
    const std::string base = std::string(dir) + "/bk_gen_" + std::to_string(serial++);
    if (FILE *f = fopen((base + ".cu").c_str(), "w")) fwrite(g->source.data(), 1, g->source.size(), f), fclose(f);
    if (FILE *f = fopen((base + ".cubin").c_str(), "wb")) fwrite(g->cubin.data(), 1, g->cubin.size(), f), fclose(f);
  }
  return g;
}

void gen_destroy(GenStencil *g) {
  if (!g) return;
  for (auto &kv : g->loaded) cudaLibraryUnload(kv.second.first);
  delete g;
}

const std::string &gen_source(const GenStencil *g) { return g->source; }
size_t gen_cubin_bytes(const GenStencil *g) { return g->cubin.size(); }

int gen_launch(GenStencil *g, const bk_field_t &f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
               const unsigned *hi, cudaStream_t s, int part, const unsigned *ready_lo, const unsigned *ready_hi) {
  int dev = 0;
  BK_CUDA(cudaGetDevice(&dev));
  cudaKernel_t kern;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    auto it = g->loaded.find(dev);
    if (it == g->loaded.end()) {
      cudaLibrary_t lib;
      BK_CUDA(cudaLibraryLoadData(&lib, g->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
      cudaKernel_t k;
      BK_CUDA(cudaLibraryGetKernel(&k, lib, "bk_gen"));
      it = g->loaded.emplace(dev, std::make_pair(lib, k)).first;
      if (getenv("BK_DEBUG")) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, (const void *) k) == cudaSuccess)
          fprintf(stderr, "[bk] generated kernel: %zu taps, %d regs, %zu B local, %zu B smem, cubin %zu B\n", g->taps.size(),
                  fa.numRegs, fa.localSizeBytes, g->geom.smem, g->cubin.size());
      }
    }
    kern = it->second.second;
  }
  return launch_generated((const void *) kern, g->geom, g->coef.data(), f, grid, gdims, lo, hi, s, part, ready_lo, ready_hi);
}

}  // namespace bk
