// api.cu -- the extern "C" boundary declared in include/mhdflows_b200.h: argument checks, error translation, dispatch
// to the Solver<T> behind the handle (solver.cuh).
#include "solver.cuh"

// ---- extern "C" --------------------------------------------------------------------------------
template <typename Fn> static int guard(mhdf_handle* h, Fn fn) {
  if (h == nullptr) return MHDF_ERR_INVALID;
  try {
    fn();
    return MHDF_OK;
  } catch (const Err& e) {
    h->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return MHDF_ERR_STATE;
  }
}

extern "C" {

int mhdf_create(const mhdf_config* c, mhdf_handle** out) {
  if (c == nullptr || out == nullptr) { g_create_error = "null argument"; return MHDF_ERR_INVALID; }
  *out = nullptr;
  auto bad = [&](const char* m) { g_create_error = m; return (int)MHDF_ERR_INVALID; };
  for (int n : {c->nx, c->ny, c->nz})
    if (!is_pow2(n) || n < 16 || n > 1024) return bad("nx, ny, nz must be powers of two in 16..1024");
  if ((long long)c->ny * c->nz < 256) return bad("ny*nz must be at least 256");
  if (!(c->Lx > 0 && c->Ly > 0 && c->Lz > 0)) return bad("Lx, Ly, Lz must be positive");
  if (c->physics < MHDF_HD || c->physics > MHDF_EMHD) return bad("physics must be MHDF_HD, MHDF_MHD or MHDF_EMHD");
  if (c->stepper != MHDF_RK4 && c->stepper != MHDF_LSRK54 && c->stepper != MHDF_HM89) return bad("stepper must be RK4, LSRK54 or HM89 (Problems.jl:123-128)");
  if (c->stepper == MHDF_HM89 && c->physics != MHDF_EMHD) return bad("the HM89 stepper exists for EMHD problems only (Problems.jl:124)");
  if (c->dtype != MHDF_F32 && c->dtype != MHDF_F64) return bad("dtype must be MHDF_F32 or MHDF_F64");
  if (c->nranks < 1 || c->nranks > 64 || c->rank < 0 || c->rank >= c->nranks) return bad("need 0 <= rank < nranks <= 64");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    return MHDF_ERR_CUDA;
  }
  if (c->device < 0 || c->device >= ndev) return bad("device ordinal out of range");
  try {
    if (c->dtype == MHDF_F32) *out = mhdf_make_solver_f32(*c);
    else *out = mhdf_make_solver_f64(*c);
    return MHDF_OK;
  } catch (const Err& er) {
    g_create_error = er.msg;
    return er.code;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return MHDF_ERR_STATE;
  }
}

int mhdf_destroy(mhdf_handle* h) {
  if (h == nullptr) return MHDF_ERR_INVALID;
  delete h;
  return MHDF_OK;
}
const char* mhdf_last_error(const mhdf_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int mhdf_nccl_unique_id(void* id128) {
  if (id128 == nullptr) return MHDF_ERR_INVALID;
  std::memset(id128, 0, 128);
  std::string why;
  if (!g_nccl.load(why)) { g_create_error = why; return MHDF_ERR_NCCL; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return MHDF_ERR_NCCL; }
  std::memcpy(id128, &id, 128);
  return MHDF_OK;
}
int mhdf_set_real(mhdf_handle* h, int f, const void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->set_real(f, p); }); }
int mhdf_get_real(mhdf_handle* h, int f, int w, void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->get_real(f, w, p); }); }
int mhdf_set_spectral(mhdf_handle* h, int f, const void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->set_spectral(f, p); }); }
int mhdf_get_spectral(mhdf_handle* h, int f, int w, void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->get_spectral(f, w, p); }); }
int mhdf_stepper_stats(mhdf_handle* h, long long* it, double* eps) { return guard(h, [&] { h->stepper_stats(it, eps); }); }
int mhdf_step(mhdf_handle* h, int n) { return guard(h, [&] { if (n < 0) throw Err{MHDF_ERR_INVALID, "nsteps < 0"}; h->step(n); }); }
int mhdf_calcN(mhdf_handle* h, void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->calcN(p); }); }
int mhdf_set_dt(mhdf_handle* h, double dt) { return guard(h, [&] { h->set_dt(dt); }); }
int mhdf_set_clock(mhdf_handle* h, double t, long long s) { return guard(h, [&] { h->set_clock(t, s); }); }
int mhdf_get_clock(const mhdf_handle* h, double* t, double* dt, long long* s) {
  if (!h) return MHDF_ERR_INVALID;
  h->get_clock(t, dt, s);
  return MHDF_OK;
}
int mhdf_cfl_dt(mhdf_handle* h, double coef, double t_diff, double* dt) { return guard(h, [&] { h->cfl_dt(coef, t_diff, dt); }); }
int mhdf_energy(mhdf_handle* h, int w, double* KE, double* ME) { return guard(h, [&] { h->energy(w, KE, ME); }); }
int mhdf_helicity(mhdf_handle* h, double* a, double* b, double* c) { return guard(h, [&] { h->helicity(a, b, c); }); }
int mhdf_spectrum(mhdf_handle* h, int f, double* Pk, int nb) { return guard(h, [&] { if (!Pk) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->spectrum(f, Pk, nb); }); }
int mhdf_stale_stats(const mhdf_handle* h, double* mx, double* sm) {
  if (!h) return MHDF_ERR_INVALID;
  h->stale_stats(mx, sm);
  return MHDF_OK;
}
int mhdf_step_timed(mhdf_handle* h, int n, double* ms) { return guard(h, [&] { if (n < 0 || !ms) throw Err{MHDF_ERR_INVALID, "bad argument"}; h->step_timed(n, ms); }); }
int mhdf_profile(mhdf_handle* h, int en) { return guard(h, [&] { h->profile(en); }); }
int mhdf_profile_get(mhdf_handle* h, double* ms, long long* cnt, int n) { return guard(h, [&] { h->profile_get(ms, cnt, n); }); }
int mhdf_set_forcing(mhdf_handle* h, int f, const void* p) { return guard(h, [&] { h->set_forcing(f, p); }); }
int mhdf_set_forcing_a99(mhdf_handle* h, const mhdf_a99* p) { return guard(h, [&] { h->set_forcing_a99(p); }); }
int mhdf_forcing_a99_calls(const mhdf_handle* h, unsigned long long* calls) {
  if (!h || !calls) return MHDF_ERR_INVALID;
  *calls = h->a99_calls();
  return MHDF_OK;
}
int mhdf_set_forcing_callback(mhdf_handle* h, mhdf_forcing_fn fn, void* user) { return guard(h, [&] { h->set_forcing_callback(fn, user); }); }
int mhdf_set_forcing_spectral(mhdf_handle* h, int field, const void* p) { return guard(h, [&] { h->set_forcing_spectral(field, p); }); }
int mhdf_set_vp_field(mhdf_handle* h, int which, const void* p) { return guard(h, [&] { if (!p) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->set_vp_field(which, p); }); }
int mhdf_set_forcing_nd(mhdf_handle* h, double P, const void* fx, const void* fy, const void* fz) { return guard(h, [&] { h->set_forcing_nd(P, fx, fy, fz); }); }
int mhdf_div_correction(mhdf_handle* h, int group) { return guard(h, [&] { h->div_correction(group); }); }
int mhdf_scale_decomposition(mhdf_handle* h, int group, int which, double k1, double k2, void* out3) {
  return guard(h, [&] { if (!out3) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->analysis(0, group, which, k1, k2, out3); });
}
int mhdf_vector_potential(mhdf_handle* h, int which, void* out3) {
  return guard(h, [&] { if (!out3) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->analysis(1, 1, which, 0, 0, out3); });
}
int mhdf_correlation(mhdf_handle* h, int group, int which, void* out3) {
  return guard(h, [&] { if (!out3) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->analysis(2, group, which, 0, 0, out3); });
}
int mhdf_set_random_phase(mhdf_handle* h, int group, unsigned long long seed, double k0, double P, double k_peak) {
  return guard(h, [&] { h->set_random_phase(group, seed, k0, P, k_peak); });
}
int mhdf_ipc_blob_size(const mhdf_handle*) { return (int)sizeof(IpcBlob); }
int mhdf_ipc_export(mhdf_handle* h, void* blob) { return guard(h, [&] { if (!blob) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->ipc_export(blob); }); }
int mhdf_ipc_import(mhdf_handle* h, const void* blobs) { return guard(h, [&] { if (!blobs) throw Err{MHDF_ERR_INVALID, "null pointer"}; h->ipc_import(blobs); }); }
long long mhdf_launch_count(const mhdf_handle* h) { return h ? h->launch_count() : -1; }
int mhdf_info(const mhdf_handle* h, int* nf, int* kx, int* kxp, int* ky, int* kz, long long* bytes) {
  if (!h) return MHDF_ERR_INVALID;
  h->info(nf, kx, kxp, ky, kz, bytes);
  return MHDF_OK;
}

}  // extern "C"
