// Driver for sanitizer runs of the WHOLE library on the CPU emulator (api.cu + solver.cuh + kernels compiled with
// -DMHDF_CPU_EMU -fsanitize=address,undefined or -fsanitize=thread and linked with this file): device buffers are heap blocks
// of exactly the sizes the solver computes, so any kernel or copy that leaves a buffer is a heap-buffer-overflow.
// Drives the C ABI only (no oracle): HD / MHD / EMHD, RK4 / LSRK54, both EMHD x-kernel forms, constant forcing,
// A99 driving, negative damping, a calcF! host callback, volume penalisation, divergence corrections, the HM89 stepper, the
// random-phase initial condition, the analysis entry points, calcN, get/set real and spectral, diagnostics, spectrum.
// Built and run by tests/test_emulated_library.py when MHDF_EMU_SANITIZE_LIB=1.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mhdflows_b200.h"

static int g_fail = 0;
#define OK(call)                                                                            \
  do {                                                                                      \
    const int rc_ = (call);                                                                 \
    if (rc_ != 0) { std::printf("FAIL %s -> %d (%s)\n", #call, rc_, mhdf_last_error(h)); ++g_fail; } \
  } while (0)

// calcF! host callback: reads the stage input of field 0 and uploads a scaled copy as the forcing of field 1
template <typename T> struct Cb { mhdf_handle* h; std::vector<T> buf; int calls; };
template <typename T> static int cb_fn(void* user, double t) {
  auto* c = static_cast<Cb<T>*>(user);
  if (mhdf_get_spectral(c->h, 0, MHDF_STAGE, c->buf.data()) != 0) return 1;
  for (auto& v : c->buf) v = (T)(0.1 * (1.0 + t)) * v;
  ++c->calls;
  return mhdf_set_forcing_spectral(c->h, 1, c->buf.data());
}

template <typename T>
static void run(const char* label, int physics, int stepper, int nx, int ny, int nz, bool vp, bool a99, bool forcing, bool nd = false, bool cbf = false) {
  mhdf_config c;
  std::memset(&c, 0, sizeof c);
  c.nx = nx; c.ny = ny; c.nz = nz; c.Lx = c.Ly = c.Lz = 2 * M_PI;
  c.nu = 1e-2; c.eta = 1e-2; c.n_nu = physics == MHDF_MHD ? 2 : 0; c.dt = physics == MHDF_EMHD ? 1e-4 : 1e-3;
  c.physics = physics; c.stepper = stepper; c.dtype = sizeof(T) == 4 ? MHDF_F32 : MHDF_F64; c.nranks = 1; c.vp = vp; c.nd = nd;
  mhdf_handle* h = nullptr;
  if (mhdf_create(&c, &h) != 0) { std::printf("FAIL create %s: %s\n", label, mhdf_last_error(nullptr)); ++g_fail; return; }
  const size_t n = (size_t)nx * ny * nz, ns = (size_t)(nx / 2 + 1) * ny * nz;
  const int F = physics == MHDF_MHD ? 6 : 3;
  std::vector<T> re(n), back(n);
  std::vector<T> spec(2 * ns * F);
  for (int f = 0; f < F; ++f) {
    for (size_t i = 0; i < n; ++i) re[i] = (T)(0.1 * std::sin(0.37 * (double)i + f) + 0.05 * std::cos(0.011 * (double)i * (f + 1)));
    OK(mhdf_set_real(h, f, re.data()));
  }
  if (forcing) OK(mhdf_set_forcing(h, 1, re.data()));
  if (nd) OK(mhdf_set_forcing_nd(h, 0.5, re.data(), re.data(), re.data()));
  Cb<T> cb{h, std::vector<T>(2 * ns), 0};
  if (cbf) OK(mhdf_set_forcing_callback(h, cb_fn<T>, &cb));
  if (a99) { mhdf_a99 q{MHDF_A99_HOST, 0.5, 2.0, 1.0, 1.0, 42ull, 0ull}; OK(mhdf_set_forcing_a99(h, &q)); }
  if (vp) for (int w = 0; w <= F; ++w) {
    for (size_t i = 0; i < n; ++i) re[i] = w == 0 ? (T)((i / 7) % 2) : (T)(0.01 * w);
    OK(mhdf_set_vp_field(h, w, re.data()));
  }
  OK(mhdf_step(h, 1));
  OK(mhdf_calcN(h, spec.data()));
  if (physics != MHDF_EMHD) OK(mhdf_div_correction(h, 0));
  if (physics != MHDF_HD) OK(mhdf_div_correction(h, 1));
  OK(mhdf_step(h, 1));
  double dt = 0, ke = 0, me = 0, hk, hm, hc;
  OK(mhdf_cfl_dt(h, 0.25, 1e-3, &dt));
  OK(mhdf_energy(h, MHDF_FRESH, &ke, &me));
  OK(mhdf_energy(h, MHDF_STALE, &ke, &me));
  OK(mhdf_helicity(h, &hk, &hm, &hc));
  std::vector<double> Pk(32);
  OK(mhdf_spectrum(h, 0, Pk.data(), 32));
  for (int f = 0; f < F; ++f) {
    OK(mhdf_get_real(h, f, MHDF_STALE, back.data()));
    OK(mhdf_get_spectral(h, f, MHDF_FRESH, spec.data()));
    OK(mhdf_set_spectral(h, f, spec.data()));
  }
  OK(mhdf_step(h, 1));
  std::vector<T> three(3 * n);
  for (int grp = (physics == MHDF_EMHD ? 1 : 0); grp <= (physics == MHDF_HD ? 0 : 1); ++grp) {
    OK(mhdf_scale_decomposition(h, grp, MHDF_STALE, 1.0, 4.0, three.data()));
    OK(mhdf_correlation(h, grp, MHDF_FRESH, three.data()));
  }
  if (physics != MHDF_HD) OK(mhdf_vector_potential(h, MHDF_FRESH, three.data()));
  long long iters = 0;
  double eps = 0;
  OK(mhdf_stepper_stats(h, &iters, &eps));
  if (stepper == MHDF_HM89 && !(iters >= 1 && eps <= 5e-4)) { std::printf("FAIL %s: HM89 iterations %lld eps %g\n", label, iters, eps); ++g_fail; }
  if (cbf && cb.calls < 8) { std::printf("FAIL %s: callback ran %d times\n", label, cb.calls); ++g_fail; }
  OK(mhdf_set_random_phase(h, physics == MHDF_EMHD ? 1 : 0, 99ull, -5.0 / 6, 1.0, 0.0));
  OK(mhdf_step(h, 1));
  const bool finite = std::isfinite(ke) && std::isfinite(me) && std::isfinite((double)back[n / 2]);
  std::printf("%s %s (KE %.3e ME %.3e, %lld launches)\n", finite ? "PASS" : "FAIL", label, ke, me, mhdf_launch_count(h));
  if (!finite) ++g_fail;
  OK(mhdf_destroy(h));
}

int main() {
  run<float>("hd rk4 16x32x16", MHDF_HD, MHDF_RK4, 16, 32, 16, false, false, false);
  run<float>("mhd rk4 32x16x16 forcing + a99", MHDF_MHD, MHDF_RK4, 32, 16, 16, false, true, true);
  run<float>("mhd lsrk54 16^3", MHDF_MHD, MHDF_LSRK54, 16, 16, 16, false, false, false);
  run<double>("mhd rk4 f64 16x16x32", MHDF_MHD, MHDF_RK4, 16, 16, 32, false, false, false);
  setenv("MHDF_EMHD2", "0", 1);
  run<float>("emhd rk4 register-form x kernel 16x16x32", MHDF_EMHD, MHDF_RK4, 16, 16, 32, false, false, false);
  setenv("MHDF_EMHD2", "1", 1);
  run<float>("emhd lsrk54 second x-kernel form 16^3", MHDF_EMHD, MHDF_LSRK54, 16, 16, 16, false, false, false);
  run<double>("emhd rk4 f64 second x-kernel form 16^3", MHDF_EMHD, MHDF_RK4, 16, 16, 16, false, false, false);
  unsetenv("MHDF_EMHD2");
  run<float>("mhd rk4 a99 16^3", MHDF_MHD, MHDF_RK4, 16, 16, 16, false, true, false);
  run<float>("hd rk4 volume penalisation 16x16x32", MHDF_HD, MHDF_RK4, 16, 16, 32, true, false, false);
  run<float>("mhd rk4 volume penalisation + a99 16^3", MHDF_MHD, MHDF_RK4, 16, 16, 16, true, true, false);
  run<double>("mhd lsrk54 f64 volume penalisation 16^3", MHDF_MHD, MHDF_LSRK54, 16, 16, 16, true, false, false);
  run<float>("emhd hm89 16x16x32", MHDF_EMHD, MHDF_HM89, 16, 16, 32, false, false, false);
  run<double>("emhd hm89 f64 16^3", MHDF_EMHD, MHDF_HM89, 16, 16, 16, false, false, false);
  run<float>("mhd rk4 negative damping 16^3", MHDF_MHD, MHDF_RK4, 16, 16, 16, false, false, false, true, false);
  run<float>("mhd lsrk54 calcF callback 16x32x16", MHDF_MHD, MHDF_LSRK54, 16, 32, 16, false, false, false, false, true);
  std::printf("library sanitize driver done: %d failure(s)\n", g_fail);
  return g_fail ? 1 : 0;
}
