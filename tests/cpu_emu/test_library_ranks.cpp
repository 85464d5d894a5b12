// Slab-decomposed runs of the WHOLE library on the CPU emulator: P ranks = P host threads of this process, each with its own
// handle; NCCL and CUDA IPC are the in-process stand-ins of cuda_host_emu.h (a collective = a rendezvous of the rank threads,
// a peer push = a memcpy into the peer's buffer).  Every stream is synchronous here, so this checks the DATA PATH of the
// distributed code -- blocked exchange layouts, piece offsets, slab bounds, the mirror-plane gather, all-reduced statistics,
// the z-chunk pipelined path (MHDF_ZCHUNKS) -- not stream / event ordering.  Expected: the spectral state of the P-rank run
// equals the single-rank run bit for bit.  Usage: test_library_ranks [P] (env MHDF_ZCHUNKS, MHDF_PEER=0 for send/recv).
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mhdflows_b200.h"

static int g_fail = 0;
static long long g_launches = 0;   // kernel launches of rank 0 in the last run (the pipelined path launches its passes per z chunk)
static void band(int n, int* lo, int* hi0) {   // FourierFlows.getaliasedwavenumbers, == solver.cuh alias_range
  const double af = 1.0 / 3.0, L = (1.0 - af) / 2.0, R = (1.0 + af) / 2.0;
  *lo = (int)std::floor(L * n) + 1 - 1;
  *hi0 = (int)std::ceil(R * n);
}

// calcF! as a host callback (mhdf_set_forcing_callback): the forcing of field 0 is a fixed spectral array scaled by (1 + 10 t); every
// rank uploads its own ky rows from inside the callback, the library brackets the callback with cross-rank barriers
template <typename T> struct CbCtx { mhdf_handle* h; std::vector<T> base, scaled; int calls; };
template <typename T> static int forcing_cb(void* user, double t) {
  auto* c = static_cast<CbCtx<T>*>(user);
  for (size_t i = 0; i < c->base.size(); ++i) c->scaled[i] = (T)((1.0 + 10.0 * t) * (double)c->base[i]);
  ++c->calls;
  return mhdf_set_forcing_spectral(c->h, 0, c->scaled.data());
}

template <typename T>
static std::vector<std::vector<T>> run(int P, bool peer, int physics, int stepper, int nx, int ny, int nz, int steps, bool a99, bool vp, bool cbf = false) {
  const int F = physics == MHDF_MHD ? 6 : 3;
  const int nkr = nx / 2 + 1;
  int lo, hi0;
  band(ny, &lo, &hi0);
  const int Ky = lo + (ny - hi0), Kyl = (Ky + P - 1) / P, nzl = nz / P;
  const size_t nreal = (size_t)nx * ny * nz;
  std::vector<std::vector<T>> fields(F, std::vector<T>(nreal)), vpf(1 + F, std::vector<T>(nreal));
  for (int f = 0; f < F; ++f)
    for (size_t i = 0; i < nreal; ++i) fields[f][i] = (T)(0.1 * std::sin(0.37 * (double)i + f) + 0.05 * std::cos(0.011 * (double)i * (f + 1)));
  for (int w = 0; w <= F; ++w)
    for (size_t i = 0; i < nreal; ++i) vpf[w][i] = w == 0 ? (T)((i / 5) % 2) : (T)(0.01 * w);
  char id[128];
  std::memset(id, 0, sizeof id);
  if (P > 1 && mhdf_nccl_unique_id(id) != 0) { std::printf("FAIL nccl id\n"); ++g_fail; return {}; }
  std::vector<std::vector<T>> full(F, std::vector<T>(2 * (size_t)nkr * ny * nz, (T)0));   // assembled (nkr, ny, nz) spectra
  std::vector<char> blobs;
  std::barrier<> sync(P);
  std::vector<double> energies(2 * P);
  auto body = [&](int r) {
    mhdf_config c;
    std::memset(&c, 0, sizeof c);
    c.nx = nx; c.ny = ny; c.nz = nz; c.Lx = c.Ly = c.Lz = 2 * M_PI;
    c.nu = 1e-2; c.eta = 2e-2; c.dt = physics == MHDF_EMHD ? 1e-4 : 2e-3; c.physics = physics; c.stepper = stepper;
    c.dtype = sizeof(T) == 4 ? MHDF_F32 : MHDF_F64; c.rank = r; c.nranks = P; c.nccl_id = P > 1 ? id : nullptr; c.vp = vp;
    mhdf_handle* h = nullptr;
    if (mhdf_create(&c, &h) != 0) { std::printf("FAIL create rank %d: %s\n", r, mhdf_last_error(nullptr)); ++g_fail; std::exit(2); }
    auto ok = [&](int rc, const char* what) { if (rc != 0) { std::printf("FAIL %s rank %d: %s\n", what, r, mhdf_last_error(h)); ++g_fail; std::exit(2); } };
    if (P > 1 && peer) {
      const int bs = mhdf_ipc_blob_size(h);
      if (r == 0) blobs.assign((size_t)bs * P, 0);
      sync.arrive_and_wait();
      ok(mhdf_ipc_export(h, blobs.data() + (size_t)r * bs), "ipc_export");
      sync.arrive_and_wait();
      ok(mhdf_ipc_import(h, blobs.data()), "ipc_import");
      sync.arrive_and_wait();
    }
    const size_t slab = (size_t)nx * ny * nzl;
    for (int f = 0; f < F; ++f) ok(mhdf_set_real(h, f, fields[f].data() + (size_t)r * slab), "set_real");
    if (vp) for (int w = 0; w <= F; ++w) ok(mhdf_set_vp_field(h, w, vpf[w].data() + (size_t)r * slab), "set_vp_field");
    if (a99) { mhdf_a99 q{MHDF_A99_HOST, 0.5, 2.0, 1.0, 1.0, 42ull, 0ull}; ok(mhdf_set_forcing_a99(h, &q), "a99"); }
    CbCtx<T> ctx{h, {}, {}, 0};
    if (cbf) {
      const int nyh = P > 1 ? Kyl : ny;
      ctx.base.assign(2 * (size_t)nkr * nyh * nz, (T)0);
      for (int k = 0; k < nz; ++k)
        for (int j = 0; j < nyh; ++j) {
          int iy = j;
          if (P > 1) { const int jg = r * Kyl + j; if (jg >= Ky) continue; iy = jg < lo ? jg : jg + (hi0 - lo); }
          for (int i = 0; i < nkr; ++i) {
            const size_t o = 2 * (((size_t)k * nyh + j) * nkr + i);
            ctx.base[o] = (T)(0.3 * std::sin(0.7 * i + 1.3 * iy + 2.1 * k));
            ctx.base[o + 1] = (T)(i == 0 ? 0.0 : 0.2 * std::cos(1.1 * i - 0.4 * iy + 0.9 * k));
          }
        }
      ctx.scaled = ctx.base;
      ok(mhdf_set_forcing_callback(h, forcing_cb<T>, &ctx), "set_forcing_callback");
    }
    ok(mhdf_step(h, steps), "step");
    if (physics != MHDF_EMHD) ok(mhdf_div_correction(h, 0), "DivVCorrection");
    ok(mhdf_step(h, 1), "step");
    ok(mhdf_energy(h, MHDF_FRESH, &energies[2 * r], &energies[2 * r + 1]), "energy");
    // local spectra: (nkr, ny, nz) on one rank, (nkr, Kyl, nz) = the rank's compact ky rows on several
    const int nyh = P > 1 ? Kyl : ny;
    std::vector<T> loc(2 * (size_t)nkr * nyh * nz);
    for (int f = 0; f < F; ++f) {
      ok(mhdf_get_spectral(h, f, MHDF_FRESH, loc.data()), "get_spectral");
      for (int k = 0; k < nz; ++k)
        for (int j = 0; j < nyh; ++j) {
          int iy = j;
          if (P > 1) {
            const int jg = r * Kyl + j;
            if (jg >= Ky) continue;
            iy = jg < lo ? jg : jg + (hi0 - lo);
          }
          std::memcpy(&full[f][2 * (((size_t)k * ny + iy) * nkr)], &loc[2 * (((size_t)k * nyh + j) * nkr)], 2 * (size_t)nkr * sizeof(T));
        }
    }
    if (cbf && ctx.calls != 4 * (steps + 1)) { std::printf("FAIL forcing callback ran %d times\n", ctx.calls); ++g_fail; }
    if (r == 0) g_launches = mhdf_launch_count(h);
    sync.arrive_and_wait();     // nobody tears its buffers down while a peer may still address them
    ok(mhdf_destroy(h), "destroy");
  };
  std::vector<std::thread> th;
  for (int r = 0; r < P; ++r) th.emplace_back(body, r);
  for (auto& t : th) t.join();
  for (int r = 1; r < P; ++r)
    if (std::fabs(energies[2 * r] - energies[0]) > 1e-12 * std::fabs(energies[0])) { std::printf("FAIL energies differ between ranks\n"); ++g_fail; }
  full.push_back(std::vector<T>{(T)energies[0], (T)energies[1]});
  return full;
}

template <typename T>
static void compare(const char* label, int P, bool peer, int physics, int stepper, int nx, int ny, int nz, bool a99, bool vp, bool cbf = false) {
  auto one = run<T>(1, false, physics, stepper, nx, ny, nz, 1, a99, vp, cbf);
  const long long l1 = g_launches;
  auto many = run<T>(P, peer, physics, stepper, nx, ny, nz, 1, a99, vp, cbf);
  const long long lP = g_launches;
  bool same = one.size() == many.size() && !one.empty();
  double norm = 0;
  for (size_t f = 0; same && f + 1 < one.size(); ++f) {
    same = std::memcmp(one[f].data(), many[f].data(), one[f].size() * sizeof(T)) == 0;
    for (T v : one[f]) norm += (double)v * v;
  }
  const double e1 = (double)one.back()[0] + (double)one.back()[1], e2 = same ? (double)many.back()[0] + (double)many.back()[1] : 0;
  const double de = same ? std::fabs(e1 - e2) / std::fabs(e1) : 1;
  const bool ok = same && norm > 0 && de < 1e-5;
  std::printf("%s ranks-vs-single %s P=%d %s (norm %.3e, energy rel diff %.1e, launches %lld -> %lld)\n", ok ? "PASS" : "FAIL", label, P, peer ? "peer pushes" : "send/recv", std::sqrt(norm), de, l1, lP);
  if (!ok) ++g_fail;
}

int main(int argc, char** argv) {
  const int P = argc > 1 ? std::atoi(argv[1]) : 2;
  const char* e = std::getenv("MHDF_PEER");
  const bool peer = !(e && std::atoi(e) == 0);
  if (argc >= 5) {   // custom shape: test_library_ranks P nx ny nz [hd|mhd|emhd]
    const int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), nz = std::atoi(argv[4]);
    const std::string ph = argc > 5 ? argv[5] : "mhd";
    const int phys = ph == "hd" ? MHDF_HD : (ph == "emhd" ? MHDF_EMHD : MHDF_MHD);
    compare<float>((ph + " rk4 custom " + std::to_string(nx) + "x" + std::to_string(ny) + "x" + std::to_string(nz)).c_str(), P, peer, phys, MHDF_RK4, nx, ny, nz, false, false);
    std::printf("library ranks driver done: %d failure(s)\n", g_fail);
    return g_fail ? 1 : 0;
  }
  const bool quick = std::getenv("MHDF_RANKS_QUICK") != nullptr;   // CI: two cases; the full set takes ~90 s per configuration
  if (!quick) compare<float>("mhd rk4 16x16x32", P, peer, MHDF_MHD, MHDF_RK4, 16, 16, 32, false, false);
  if (!quick) compare<float>("hd lsrk54 16x32x16", P, peer, MHDF_HD, MHDF_LSRK54, 16, 32, 16, false, false);
  compare<double>("emhd rk4 f64 16x16x16", P, peer, MHDF_EMHD, MHDF_RK4, 16, 16, 16, false, false);
  compare<float>("mhd rk4 a99 + vp 16x16x16", P, peer, MHDF_MHD, MHDF_RK4, 16, 16, 16, true, true);
  compare<double>("emhd hm89 f64 16x16x16", P, peer, MHDF_EMHD, MHDF_HM89, 16, 16, 16, false, false);   // fixed-point loop: max over ranks
  compare<float>("mhd rk4 calcF callback 16x16x16", P, peer, MHDF_MHD, MHDF_RK4, 16, 16, 16, false, false, true);   // API traffic between the stages of a step
  std::printf("library ranks driver done: %d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
