// solver.cuh -- handle and orchestration behind the extern "C" boundary (api.cu) declared in include/mhdflows_b200.h.
// Solver<float> and Solver<double> are instantiated in their own translation units (solver_f32.cu, solver_f64.cu) so the
// two precisions compile in parallel.
//
// One RHS evaluation (reference: MHDcalcN!/HDcalcN!/EMHDcalcN!, src/pgen.jl:153-181) is
//   [EMHD: derive] -> inverse z pass -> inverse y pass -> fused x pass (c2r, products, r2c)
//   -> forward y pass -> forward z pass -> spectral assembly + Runge-Kutta stage update
// on a compact state that stores only the modes FourierFlows' dealias!() keeps.
#pragma once
#ifdef MHDF_CPU_EMU
#include "cuda_host_emu.h"   // tests/cpu_emu: the library compiled as plain C++ for the CPU test-suite (never shipped)
#define MHDF_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch_call(dim3(grid), (int)(block), [&] { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#include <nccl.h>
// one spelling for every kernel launch (the kernel name is parenthesised because template argument lists contain commas)
#define MHDF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mhdflows_b200.h"
#include "kernels.cuh"

using namespace mhdf;

inline thread_local std::string g_create_error;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      throw Err{MHDF_ERR_CUDA, buf_};                                                              \
    }                                                                                              \
  } while (0)

struct Err {
  int code;
  std::string msg;
};

// NCCL is bound at run time (dlopen) so the library has no link-time dependency on a particular libnccl; when
// the host process already loaded one (e.g. PyTorch's) that copy is used.
struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& why) {
    if (so) return true;
#ifdef MHDF_CPU_EMU   // CPU test-suite: the in-process stand-ins of tests/cpu_emu/cuda_host_emu.h (ranks = threads)
    (void)why;
    so = this;
    GetUniqueId = emu_nccl::GetUniqueId; CommInitRank = emu_nccl::CommInitRank; CommDestroy = emu_nccl::CommDestroy;
    Send = emu_nccl::Send; Recv = emu_nccl::Recv; GroupStart = emu_nccl::GroupStart; GroupEnd = emu_nccl::GroupEnd;
    AllReduce = emu_nccl::AllReduce; AllGather = emu_nccl::AllGather;
    GetErrorString = emu_nccl::GetErrorString;
    return true;
#endif
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
    if (!so) { why = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define LD(f) f = reinterpret_cast<decltype(f)>(dlsym(so, "nccl" #f)); if (!f) { why = "libnccl lacks nccl" #f; return false; }
    LD(GetUniqueId) LD(CommInitRank) LD(CommDestroy) LD(Send) LD(Recv) LD(GroupStart) LD(GroupEnd) LD(AllReduce) LD(AllGather)
    LD(GetErrorString)
#undef LD
    return true;
  }
};
inline NcclApi g_nccl;

#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
      throw Err{MHDF_ERR_NCCL, buf_};                                                              \
    }                                                                                              \
  } while (0)

enum { KC_ZINV = 0, KC_YINV, KC_XFUSED, KC_YFWD, KC_ZFWD, KC_SPEC, KC_DERIVE, KC_EXCH, KC_COUNT };

struct mhdf_handle {
  std::string err;
  virtual ~mhdf_handle() {}
  virtual void set_real(int field, const void* p) = 0;
  virtual void get_real(int field, int which, void* p) = 0;
  virtual void set_spectral(int field, const void* p) = 0;
  virtual void get_spectral(int field, int which, void* p) = 0;
  virtual void step(int n) = 0;
  virtual void calcN(void* p) = 0;
  virtual void stepper_stats(long long* iters, double* eps) const = 0;
  virtual void set_dt(double dt) = 0;
  virtual void set_clock(double t, long long step) = 0;
  virtual void get_clock(double* t, double* dt, long long* step) const = 0;
  virtual void cfl_dt(double coef, double t_diff, double* dt) = 0;
  virtual void energy(int which, double* KE, double* ME) = 0;
  virtual void helicity(double* Hk, double* Hm, double* Hc) = 0;
  virtual void spectrum(int field, double* Pk, int nbins) = 0;
  virtual void stale_stats(double* mx, double* sm) const = 0;
  virtual void step_timed(int n, double* ms) = 0;
  virtual void profile(int enable) = 0;
  virtual void profile_get(double* ms, long long* cnt, int n) = 0;
  virtual long long launch_count() const = 0;
  virtual void info(int* nf, int* kx, int* kxp, int* ky, int* kz, long long* bytes) const = 0;
  virtual void set_forcing(int field, const void* p) = 0;
  virtual void set_forcing_spectral(int field, const void* p) = 0;
  virtual void set_forcing_callback(mhdf_forcing_fn fn, void* user) = 0;
  virtual void set_forcing_a99(const mhdf_a99* p) = 0;
  virtual unsigned long long a99_calls() const = 0;
  virtual void div_correction(int group) = 0;
  virtual void analysis(int mode, int group, int which, double k1, double k2, void* out3) = 0;
  virtual void set_random_phase(int group, unsigned long long seed, double k0, double P, double k_peak) = 0;
  virtual void set_vp_field(int which, const void* p) = 0;
  virtual void set_forcing_nd(double P, const void* fx, const void* fy, const void* fz) = 0;
  virtual void ipc_export(void* blob) = 0;
  virtual void ipc_import(const void* blobs) = 0;
};
struct IpcBlob { cudaIpcMemHandle_t r, q, f; int device; int pad[15]; };

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// FourierFlows.getaliasedwavenumbers with aliased_fraction = 1/3, evaluated in Float64 exactly like
// the Julia expression: 1-based inclusive [iL, iR].
static void alias_range(int nk, int* iL, int* iR) {
  const double af = 1.0 / 3.0;
  const double L = (1.0 - af) / 2.0, R = (1.0 + af) / 2.0;
  *iL = (int)std::floor(L * nk) + 1;
  *iR = (int)std::ceil(R * nk);
}

template <typename T>
struct Solver : mhdf_handle {
  using C = Cx<T>;
  mhdf_config cfg;
  int nx, ny, nz, nkr;
  int Kx, Kxp, Ky, Kz;
  Band by, bz;
  // slab decomposition: P_ ranks; real space split along z (nzl planes each), spectral space along compact ky rows
  // (Kyl rows each, the last slab zero-padded).  One GPU: P_ = 1, nzl = nz, Kyl = Ky.
  int P_ = 1, rank_ = 0, nzl, Kyl, ky0;
  ncclComm_t comm = nullptr;
  Cx<T>* plane_loc = nullptr;   // kr = 0 plane of the stage input, local  [F][Kz][Kyl]
  Cx<T>* plane_all = nullptr;   // gathered                                  [P][F][Kz][Kyl]
  int phys, F, nin, nout;
  long long cf;   // elements of one compact field
  cudaStream_t st = nullptr;    // compute stream
  cudaStream_t sc = nullptr;    // communication stream (every NCCL call is issued here, ordered with events)
  std::vector<cudaEvent_t> dep_ev;
  size_t dep_next = 0;
  // peer-memory exchange (after mhdf_ipc_import): peers' R and Q buffers mapped into this process
  bool ipc_on = false;
  std::vector<C*> peerR, peerQ;
  static constexpr int NCS_MAX = 8;
  int NCS = [] { const char* e = getenv("MHDF_COPY_STREAMS"); int n = e ? atoi(e) : 7; return n < 1 ? 1 : (n > NCS_MAX ? NCS_MAX : n); }();
  cudaStream_t cs[NCS_MAX] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // copy streams (copy engines, no SMs)
  float* bar_d = nullptr;
  // cross-rank flags (kernels.cuh: k_flag_set / k_flag_wait): word [slot][sender] of the receiver's array holds the epoch of the
  // last exchange of that slot whose piece from `sender` has landed; slot = direction * 32 + z chunk
  static constexpr int FSLOTS = 64;
  unsigned* flags_d = nullptr;
  std::vector<unsigned*> peerF;
  unsigned epoch_[FSLOTS] = {};
  int* ferr_h = nullptr;   // host-mapped: set by a wait that timed out
  bool use_flags = [] { const char* e = getenv("MHDF_FLAGS"); return !e || atoi(e) != 0; }();
  // state registers (compact, F fields each)
  static constexpr int NREG_MAX = 5;
  C* reg[NREG_MAX] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int nreg_() const { return cfg.stepper == MHDF_RK4 ? 4 : (cfg.stepper == MHDF_HM89 ? 5 : 3); }   // HM89: sol, F0, F1, B0, B1
  int iY = 0;       // register holding sol
  int iStale = -1;  // register holding the last stage input (RK4), -1 if none
  C *P = nullptr, *Q = nullptr, *R = nullptr, *D = nullptr;
  size_t szP = 0, szQ = 0, szR = 0, szD = 0;
  T* bst = nullptr;   // EMHD stale real b [3][nz][ny][nx]
  C* force = nullptr; // constant spectral forcing [F][compact] (calcF! hook)
  unsigned fmask = 0;
  bool vp_on = false;  // volume-penalisation method (Problem(...; VP_method = true))
  T* vp_d = nullptr;   // [1 + F][nzl][ny][nx] real: chi, U0x, U0y, U0z[, B0x, B0y, B0z]  (params.χ, params.U₀x ...)
  bool nd_on = false;  // NDForceDriving! (pgen/NegativeDamping.jl): calcF of an MHD problem created with cfg.nd
  T* nd_d = nullptr;   // [3][nzl][ny][nx] real: usr_vars.fx, fy, fz
  double nd_P = 0;     // usr_vars.P
  A99Args<T> a99{};   // random driving (A99ForceDriving!), variant = A99_OFF: none
  unsigned long long a99_call = 0;   // forcing evaluations so far = the Philox counter word
  C *twx = nullptr, *twy = nullptr, *twz = nullptr;
  T *kxv = nullptr, *kyv = nullptr, *kzv = nullptr;
  XRed* red_d = nullptr;
  XRed* red_h = nullptr;   // pinned
  double* diag_d = nullptr;
  double* diag_h = nullptr;  // pinned
  double* spec_d = nullptr;
  int spec_cap = 0;
  // stale vars statistics (getCFL!/ProbDiagnostic read vars.* of the last RHS evaluation)
  double st_max[6] = {0, 0, 0, 0, 0, 0}, st_sum[6] = {0, 0, 0, 0, 0, 0}, st_cross = 0;
  T t_, dt_;
  long long step_ = 0;
  long long launches = 0;
  long long bytes_dev = 0;
  int nsm = 148;
  // profiling
  bool prof = false;
  struct Ev { cudaEvent_t a, b; int cls; };
  std::vector<Ev> evs;
  std::vector<Ev> ev_free;
  double prof_ms[KC_COUNT];
  long long prof_cnt[KC_COUNT];

  template <typename U> U* dalloc(size_t n) {
    U* p = nullptr;
    CK(cudaMalloc(&p, n * sizeof(U)));
    CK(cudaMemsetAsync(p, 0, n * sizeof(U), st));
    bytes_dev += (long long)(n * sizeof(U));
    return p;
  }

  explicit Solver(const mhdf_config& c) : cfg(c) {
    try {
      init(c);
    } catch (...) {   // e.g. cudaMalloc failed half-way: release what exists before reporting
      release();
      throw;
    }
  }
  void init(const mhdf_config& c) {
    nx = c.nx; ny = c.ny; nz = c.nz;
    nkr = nx / 2 + 1;
    int iL, iR;
    alias_range(nx, &iL, &iR);
    Kx = iL - 1;
    Kxp = (Kx + 7) / 8 * 8;
    alias_range(ny, &iL, &iR);
    by.n = ny; by.lo = iL - 1; by.hi0 = iR;
    alias_range(nz, &iL, &iR);
    bz.n = nz; bz.lo = iL - 1; bz.hi0 = iR;
    Ky = by.count(); Kz = bz.count();
    P_ = c.nranks; rank_ = c.rank;
    nzl = nz / P_;
    Kyl = (Ky + P_ - 1) / P_;
    ky0 = rank_ * Kyl;
    if (P_ > 1 && (nz % P_ != 0 || (P_ - 1) * Kyl >= Ky))
      throw Err{MHDF_ERR_INVALID, "grid too small for this many ranks (need nz % nranks == 0 and a non-empty ky slab per rank)"};
    // the blocked exchange addressing divides row indices by nzl / Kyl with a 32-bit multiply-high (magic = ceil(2^32 / d)),
    // which cannot represent d = 1 (found on the CPU emulator with ranks as threads, tests/cpu_emu/test_library_ranks.cpp)
    if (P_ > 1 && (nzl < 2 || Kyl < 2))
      throw Err{MHDF_ERR_INVALID, "grid too small for this many ranks (need at least 2 z planes and 2 retained ky rows per rank)"};
    phys = c.physics;
    F = (phys == MHDF_MHD) ? 6 : 3;
    nin = (phys == MHDF_MHD) ? 6 : (phys == MHDF_HD ? 3 : 18);
    nout = (phys == MHDF_MHD) ? 9 : (phys == MHDF_HD ? 6 : 3);
    vp_on = c.vp != 0;
    if (vp_on && phys == MHDF_EMHD) throw Err{MHDF_ERR_INVALID, "VP_method: the EMHD equation has no volume-penalisation terms (MHDSolver.jl:183-270)"};
    if (vp_on) nout += F;   // the penalisation products chi/eta (f_j - W_j) ride along the forward transforms
    // NDForceDriving! acts on the MHD path only (HDcalcN! loses its forcing, EMHDcalcN! never calls it: pgen.jl:164-181)
    nd_on = c.nd != 0 && phys == MHDF_MHD;
    if (nd_on && vp_on) throw Err{MHDF_ERR_INVALID, "NDForceDriving! together with VP_method is not supported on this path"};
    if (nd_on) nout += 3;   // the products f_i u_i ride along the forward transforms
    cf = (long long)Kxp * Kyl * Kz;
    t_ = (T)0; dt_ = (T)c.dt;
    for (int i = 0; i < KC_COUNT; ++i) { prof_ms[i] = 0; prof_cnt[i] = 0; }

    CK(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c.device));
    nsm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (P_ > 1) {
      CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
      dep_ev.resize(8192);   // cyclic pool; far more than one RHS evaluation can take between a record and its wait
      CK(cudaEventCreateWithFlags(&red_own, cudaEventDisableTiming));
      for (auto& e : dep_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      std::string why;
      if (!g_nccl.load(why)) throw Err{MHDF_ERR_NCCL, why};
      if (c.nccl_id == nullptr) throw Err{MHDF_ERR_INVALID, "nranks > 1 needs nccl_id (mhdf_nccl_unique_id on rank 0)"};
      ncclUniqueId id;
      std::memcpy(&id, c.nccl_id, sizeof id);
      NK(g_nccl.CommInitRank(&comm, P_, id, rank_));
    }
    const int nreg = nreg_();
    for (int i = 0; i < nreg; ++i) reg[i] = dalloc<C>((size_t)F * cf);
    // work buffers (elements).  P: inverse-z output / forward-y output (send layouts);  Q: inverse-y output (x input)
    // / second exchange target;  R: first exchange target / x output / forward-z output / API staging.
    const size_t e_zK = (size_t)nz * Kyl * Kxp;          // one field after a z pass      [nz][Kyl][Kxp]  (== P_ blocks)
    const size_t e_zy = (size_t)nzl * ny * Kxp;          // one field in x-pass layout    [nzl][ny][Kxp]
    const size_t need_stage = ((size_t)nkr * Kyl * nz > (size_t)nx * ny * nzl / 2 ? (size_t)nkr * Kyl * nz : (size_t)nx * ny * nzl / 2) + 16;
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    szP = mx((size_t)nin * e_zK, (size_t)nout * e_zK);
    szQ = mx((size_t)nin * e_zy, (size_t)nout * e_zK);
    szR = mx(mx((size_t)nin * e_zK, (size_t)nout * e_zy), mx((size_t)nout * (size_t)cf, need_stage));
    if (P_ == 1) szQ = mx(szQ, (size_t)nout * (size_t)cf);
    P = dalloc<C>(szP); Q = dalloc<C>(szQ); R = dalloc<C>(szR);
    if (phys == MHDF_EMHD) {
      szD = (size_t)18 * cf;
      D = dalloc<C>(szD);
      bst = dalloc<T>((size_t)3 * nx * ny * nzl);
    }
    if (vp_on) vp_d = dalloc<T>((size_t)(1 + F) * nx * ny * nzl);   // zero: chi = 0 means "no solid anywhere"
    if (nd_on) nd_d = dalloc<T>((size_t)3 * nx * ny * nzl);
    twx = make_tw(nx); twy = make_tw(ny); twz = make_tw(nz);
    // wavenumbers: built in Float64 then converted to T (FourierFlows ThreeDGrid; mirror utils/utils.jl:60-64)
    std::vector<T> hx(Kx), hy(Kyl), hz(Kz);
    for (int i = 0; i < Kx; ++i) hx[i] = (T)(i * (2.0 * M_PI / c.Lx));
    for (int j = 0; j < Kyl; ++j) hy[j] = (ky0 + j < Ky) ? (T)(by.wave(ky0 + j) * (2.0 * M_PI / c.Ly)) : (T)0;
    for (int k = 0; k < Kz; ++k) hz[k] = (T)(bz.wave(k) * (2.0 * M_PI / c.Lz));
    kxv = dalloc<T>(Kx); kyv = dalloc<T>(Kyl); kzv = dalloc<T>(Kz);
    CK(cudaMemcpyAsync(kxv, hx.data(), Kx * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kyv, hy.data(), Kyl * sizeof(T), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kzv, hz.data(), Kz * sizeof(T), cudaMemcpyHostToDevice, st));
    if (P_ > 1) build_tables();
    choose_chunks();
    red_d = dalloc<XRed>(1);
    diag_d = dalloc<double>(8);
    CK(cudaMallocHost(&red_h, sizeof(XRed)));
    CK(cudaMallocHost(&diag_h, 8 * sizeof(double)));
    std::memset(red_h, 0, sizeof(XRed));
    CK(cudaStreamSynchronize(st));
    setup_attrs();
  }

  // Exchange buffers are [peer][field][z'][ky'][kx]: the piece for / from one peer is contiguous, blk(n) elements for an
  // n-field batch.
  size_t blk(int nf) const { return (size_t)nf * nzl * Kyl * Kxp; }
  void build_tables() {   // slab runs: mirror-plane buffers, exchange-size check
    plane_loc = dalloc<C>((size_t)F * Kz * Kyl);
    plane_all = dalloc<C>((size_t)P_ * F * Kz * Kyl);
    flags_d = dalloc<unsigned>((size_t)FSLOTS * P_);
    CK(cudaMallocHost(&ferr_h, sizeof(int)));
    *ferr_h = 0;
    check_blk(CHUNK > 0 ? CHUNK : (nin > nout ? nin : nout));
  }
  void check_blk(int nf) const {
    if ((long long)P_ * (long long)blk(nf) >= (1LL << 31)) throw Err{MHDF_ERR_INVALID, "slab exchange buffer exceeds 2^31 elements per batch"};
  }
  ~Solver() override { release(); }
  void release() {
    cudaSetDevice(cfg.device);
    if (st) cudaStreamSynchronize(st);
    if (sc) cudaStreamSynchronize(sc);
    for (int i = 0; i < NCS; ++i) if (cs[i]) { cudaStreamSynchronize(cs[i]); cudaStreamDestroy(cs[i]); }
    if (ipc_on) {
      // peers may still be pushing into our buffers: close the mappings only after a final cross-rank barrier
      if (comm) { g_nccl.AllReduce(bar_d, bar_d, 1, ncclFloat32, ncclSum, comm, sc); cudaStreamSynchronize(sc); }
      for (int q = 0; q < P_; ++q) if (q != rank_) { cudaIpcCloseMemHandle(peerR[q]); cudaIpcCloseMemHandle(peerQ[q]); cudaIpcCloseMemHandle(peerF[q]); }
    }
    cudaFree(bar_d); cudaFree(flags_d);
    if (ferr_h) cudaFreeHost(ferr_h);
    flags_d = nullptr; ferr_h = nullptr;
    for (auto& e : dep_ev) cudaEventDestroy(e);
    if (red_own) cudaEventDestroy(red_own);
    red_own = red_ev = nullptr;
    if (sc) cudaStreamDestroy(sc);
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : ev_free) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (int i = 0; i < NREG_MAX; ++i) cudaFree(reg[i]);
    cudaFree(P); cudaFree(Q); cudaFree(R); cudaFree(D); cudaFree(bst); cudaFree(force); cudaFree(vp_d); cudaFree(nd_d); cudaFree(Xin); cudaFree(Xout); cudaFree(P2);
    cudaFree(twx); cudaFree(twy); cudaFree(twz);
    cudaFree(kxv); cudaFree(kyv); cudaFree(kzv);
    cudaFree(plane_loc); cudaFree(plane_all);
    if (comm) g_nccl.CommDestroy(comm);
    cudaFree(red_d); cudaFree(diag_d); cudaFree(spec_d);
    if (red_h) cudaFreeHost(red_h);
    if (diag_h) cudaFreeHost(diag_h);
    if (st) cudaStreamDestroy(st);
    st = sc = nullptr; comm = nullptr; ipc_on = false; red_h = nullptr; diag_h = nullptr;
    for (int i = 0; i < NREG_MAX; ++i) reg[i] = nullptr;
    P = Q = R = D = nullptr; bst = nullptr; force = nullptr; vp_d = nullptr; nd_d = nullptr; Xin = Xout = P2 = nullptr; twx = twy = twz = nullptr; kxv = kyv = kzv = nullptr;
    red_d = nullptr; diag_d = nullptr; spec_d = nullptr; plane_loc = plane_all = nullptr; bar_d = nullptr;
    dep_ev.clear(); evs.clear(); ev_free.clear();
    for (int i = 0; i < NCS_MAX; ++i) cs[i] = nullptr;
  }

  C* make_tw(int n) {
    std::vector<C> h(n);
    for (int i = 0; i < n; ++i) {
      const double a = 2.0 * M_PI * (double)i / (double)n;
      h[i].x = (T)std::cos(a);
      h[i].y = (T)(-std::sin(a));
    }
    C* d = dalloc<C>(n);
    CK(cudaMemcpyAsync(d, h.data(), n * sizeof(C), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return d;
  }

  SpecGeom<T> geom() const {
    SpecGeom<T> g;
    g.Kx = Kx; g.Kxp = Kxp; g.by = by; g.bz = bz; g.Kyl = Kyl; g.ky0 = ky0; g.F = F;
    g.kx = kxv; g.ky = kyv; g.kz = kzv; g.field = cf;
    g.mirror = (P_ > 1) ? plane_all : nullptr;
    return g;
  }

  // ---- profiling brackets ------------------------------------------------------------------
  // `b` waits for everything enqueued so far on `a`
  void order(cudaStream_t a, cudaStream_t b) {
    cudaEvent_t e = dep_ev[dep_next++ % dep_ev.size()];
    CK(cudaEventRecord(e, a));
    CK(cudaStreamWaitEvent(b, e, 0));
  }
  void sync_all() {
    CK(cudaStreamSynchronize(st));
    if (sc) CK(cudaStreamSynchronize(sc));
    for (int i = 0; i < NCS; ++i) if (cs[i]) CK(cudaStreamSynchronize(cs[i]));   // my own pushes out of the send buffers
  }
  void prof_begin(int cls, cudaStream_t s = nullptr) {
    if (!prof) return;
    if (s == nullptr) s = st;
    Ev e;
    if (!ev_free.empty()) { e = ev_free.back(); ev_free.pop_back(); }
    else { CK(cudaEventCreate(&e.a)); CK(cudaEventCreate(&e.b)); }
    e.cls = cls;
    CK(cudaEventRecord(e.a, s));
    evs.push_back(e);
  }
  void prof_end(cudaStream_t s = nullptr) {
    if (!prof) return;
    CK(cudaEventRecord(evs.back().b, s ? s : st));
    if (evs.size() > 4096) prof_collect();
  }
  void prof_collect() {
    sync_all();
    for (auto& e : evs) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e.a, e.b));
      prof_ms[e.cls] += ms;
      prof_cnt[e.cls] += 1;
      ev_free.push_back(e);
    }
    evs.clear();
  }

  // ---- kernel dispatch ------------------------------------------------------------------------
  // points per thread of the strided passes: 16 (8, 4 on short axes); 32 in the Float32 512- and 1024-point passes -- radix 32 x 32
  // (32 x 16), ONE shared-memory exchange instead of two, 128 registers, 256 / 512 threads per block.  Measured on B200, ms per
  // 1024^3 step (profiles/r02_c16_time1024.log): y inverse 26.4 -> 24.5, y forward 39.0 -> 37.4, z inverse 20.9 -> 19.1, z forward
  // 28.1 -> 24.9, step 222.1 -> 213.9; per 512^3 step (r02_c17_time1024.log): the four passes 11.7 -> 10.7, step 24.7 -> 23.6-24.1.
  // MHDF_PASS_E32: 0 = 16 points everywhere (A/B partner), 1 = the 1024-point y passes only, 2 = 1024-point y and z passes, 3 = the
  // 512-point passes as well (default).
#ifndef MHDF_PASS_E32
#define MHDF_PASS_E32 3
#endif
  static constexpr int passE(int N, bool zpass = false) {
    if (sizeof(T) == 4 && MHDF_PASS_E32 >= 1 && N >= 1024 && (!zpass || MHDF_PASS_E32 >= 2)) return 32;
    if (sizeof(T) == 4 && MHDF_PASS_E32 >= 3 && N >= 512) return 32;
    return N >= 128 ? 16 : (N >= 32 ? 8 : 4);
  }
  // columns per block: 16 (128-byte row segments); 8 in the 1024-point y passes so that two 512-thread blocks fit per SM.
  // The 1024-point z passes take 16 columns (one 1024-thread block per SM): their rows are a whole [ky][kx] plane apart, and
  // 128-byte instead of 64-byte row segments are worth more there than the second resident block -- z inverse 24.3 -> 20.8 ms,
  // z forward 37.5 -> 28.0 ms per 1024^3 step (profiles/r02_c12_time1024.log); the y passes (rows 2.7 KB apart) keep 8 columns
  // (MHDF_Y_TX16: 16 there as well -- measured slower, y inverse 26.7 -> 30.6 ms, y forward 39.2 -> 41.9 ms, r02_c13_time1024.log).
  static constexpr int passTX(int N, bool zpass = false) {
#ifdef MHDF_Y_TX16
    zpass = true;
#endif
    return sizeof(T) == 4 ? ((N >= 1024 && !zpass) ? 8 : 16) : (N >= 1024 ? 4 : 8);
  }
  static constexpr int xE(int) { return 8; }
  static constexpr int XNT = 64;   // threads per block of the x kernels: small blocks, rows decoupled per warp
  static constexpr int xRB(int N) { return XNT / (N / 2 / 8) > 0 ? XNT / (N / 2 / 8) : 1; }

  bool blk_out = false;
  template <int N, int DIR, bool ZP> void launch_pass_n(PassArgs<T>& a, int n_outer, int n_fields) {
    constexpr int E = passE(N, ZP), TX = passTX(N, ZP), R1 = imin(E, N);
    constexpr size_t smem = (size_t)PassIdx<N, TX, R1, C>::SIZE * sizeof(C);
    dim3 grid((a.inner + TX - 1) / TX, n_outer, n_fields);
    // the blocked side is the z side of the z passes and the ky side of the y passes: output of inverse-z / forward-y,
    // input of inverse-y / forward-z
    if (a.blk_rows == 0) MHDF_LAUNCH((k_pass<T, N, E, TX, DIR, (DIR > 0), 0>), grid, (N / E) * TX, smem, st, a);
    else if (a.blk2_rows > 0 && blk_out) MHDF_LAUNCH((k_pass<T, N, E, TX, DIR, (DIR > 0), 4>), grid, (N / E) * TX, smem, st, a);
    else if (a.blk2_rows > 0) MHDF_LAUNCH((k_pass<T, N, E, TX, DIR, (DIR > 0), 3>), grid, (N / E) * TX, smem, st, a);
    else if (blk_out) MHDF_LAUNCH((k_pass<T, N, E, TX, DIR, (DIR > 0), 2>), grid, (N / E) * TX, smem, st, a);
    else MHDF_LAUNCH((k_pass<T, N, E, TX, DIR, (DIR > 0), 1>), grid, (N / E) * TX, smem, st, a);
    ++launches;
  }
  template <int DIR, bool ZP = false> void launch_pass(int N, PassArgs<T>& a, int n_outer, int n_fields) {
    switch (N) {
      case 16: launch_pass_n<16, DIR, ZP>(a, n_outer, n_fields); break;
      case 32: launch_pass_n<32, DIR, ZP>(a, n_outer, n_fields); break;
      case 64: launch_pass_n<64, DIR, ZP>(a, n_outer, n_fields); break;
      case 128: launch_pass_n<128, DIR, ZP>(a, n_outer, n_fields); break;
      case 256: launch_pass_n<256, DIR, ZP>(a, n_outer, n_fields); break;
      case 512: launch_pass_n<512, DIR, ZP>(a, n_outer, n_fields); break;
      case 1024: launch_pass_n<1024, DIR, ZP>(a, n_outer, n_fields); break;
      default: throw Err{MHDF_ERR_INVALID, "unsupported axis length"};
    }
    CK(cudaGetLastError());
  }

  template <int N> static size_t x_smem() {
    constexpr int E = xE(N), M = N / 2, R1 = imin(E, M);
    return (size_t)2 * xRB(N) * RowIdx<M, R1>::SIZE * sizeof(C);
  }
  int x_grid(long long rows, int RB) const {
    long long sets = rows / RB;
    long long g = (long long)nsm * 16;
    return (int)(sets < g ? sets : g);
  }
  template <int N> void launch_xfused_n(XArgs<T>& a) {
    constexpr int E = xE(N), RB = xRB(N);
    const int grid = x_grid(a.rows, RB);
    const int threads = (N / 2 / E) * RB;
    const bool red = a.red != nullptr;
#define XLAUNCH(PH)                                                                          \
    do {                                                                                     \
      if (red) MHDF_LAUNCH((k_xfused<T, N, E, RB, PH, true>), grid, threads, x_smem<N>(), st, a);       \
      else MHDF_LAUNCH((k_xfused<T, N, E, RB, PH, false>), grid, threads, x_smem<N>(), st, a);          \
    } while (0)
    if (phys == MHDF_EMHD && emhd2) {   // opt-in second form of the EMHD kernel (MHDF_EMHD2=1): bit-identical results
      const size_t smem = x_smem<N>() + (size_t)RB * 6 * (N / 2) * sizeof(C);
      if (red) MHDF_LAUNCH((k_xfused_emhd2<T, N, E, RB, true>), grid, threads, smem, st, a);
      else MHDF_LAUNCH((k_xfused_emhd2<T, N, E, RB, false>), grid, threads, smem, st, a);
    }
    else if (nd_on && a.vp != nullptr) MHDF_LAUNCH((k_xfused<T, N, E, RB, PHYS_MHD, true, 2>), grid, threads, x_smem<N>(), st, a);
    else if (vp_on && a.vp != nullptr) {   // penalised runs: one instantiation per physics (reductions always compiled in)
      if (phys == MHDF_MHD) MHDF_LAUNCH((k_xfused<T, N, E, RB, PHYS_MHD, true, true>), grid, threads, x_smem<N>(), st, a);
      else MHDF_LAUNCH((k_xfused<T, N, E, RB, PHYS_HD, true, true>), grid, threads, x_smem<N>(), st, a);
    }
    else if (phys == MHDF_MHD) XLAUNCH(PHYS_MHD);
    else if (phys == MHDF_HD) XLAUNCH(PHYS_HD);
    else XLAUNCH(PHYS_EMHD);
#undef XLAUNCH
    ++launches;
  }
  template <int N, int DIR> void launch_xplain_n(XArgs<T>& a) {
    constexpr int E = xE(N), RB = xRB(N);
    MHDF_LAUNCH((k_xplain<T, N, E, RB, DIR>), x_grid(a.rows, RB), (N / 2 / E) * RB, x_smem<N>(), st, a);
    ++launches;
  }
  void launch_xfused(XArgs<T>& a) {
    switch (nx) {
      case 16: launch_xfused_n<16>(a); break;
      case 32: launch_xfused_n<32>(a); break;
      case 64: launch_xfused_n<64>(a); break;
      case 128: launch_xfused_n<128>(a); break;
      case 256: launch_xfused_n<256>(a); break;
      case 512: launch_xfused_n<512>(a); break;
      case 1024: launch_xfused_n<1024>(a); break;
      default: throw Err{MHDF_ERR_INVALID, "unsupported nx"};
    }
    CK(cudaGetLastError());
  }
  template <int DIR> void launch_xplain(XArgs<T>& a) {
    switch (nx) {
      case 16: launch_xplain_n<16, DIR>(a); break;
      case 32: launch_xplain_n<32, DIR>(a); break;
      case 64: launch_xplain_n<64, DIR>(a); break;
      case 128: launch_xplain_n<128, DIR>(a); break;
      case 256: launch_xplain_n<256, DIR>(a); break;
      case 512: launch_xplain_n<512, DIR>(a); break;
      case 1024: launch_xplain_n<1024, DIR>(a); break;
      default: throw Err{MHDF_ERR_INVALID, "unsupported nx"};
    }
    CK(cudaGetLastError());
  }
  void setup_attrs() {
    // opt in to > 48 KB dynamic shared memory where a plan needs it
    set_pass_attr<256>(); set_pass_attr<512>(); set_pass_attr<1024>();
    set_x_attr<16>(); set_x_attr<32>(); set_x_attr<64>(); set_x_attr<128>(); set_x_attr<256>(); set_x_attr<512>(); set_x_attr<1024>();
  }
  template <int N> void set_x_attr() {
    constexpr int E = xE(N), RB = xRB(N);
    const int smem = (int)x_smem<N>();
    if (smem > 48 * 1024) {
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_MHD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_EMHD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_MHD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_EMHD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int smem2 = smem + (int)((size_t)RB * 6 * (N / 2) * sizeof(C));
    if (smem2 > 48 * 1024) {
      CK(cudaFuncSetAttribute(k_xfused_emhd2<T, N, E, RB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      CK(cudaFuncSetAttribute(k_xfused_emhd2<T, N, E, RB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    }
    if (smem > 48 * 1024) {
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_HD, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_MHD, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xfused<T, N, E, RB, PHYS_MHD, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xplain<T, N, E, RB, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_xplain<T, N, E, RB, +1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
  }
  template <int N> void set_pass_attr() {
    set_pass_attr_tx<N, passTX(N, false), false>();
    if constexpr (passTX(N, true) != passTX(N, false) || passE(N, true) != passE(N, false)) set_pass_attr_tx<N, passTX(N, true), true>();
  }
  template <int N, int TX, bool ZP> void set_pass_attr_tx() {
    constexpr int E = passE(N, ZP), R1 = imin(E, N);
    constexpr int smem = (int)(PassIdx<N, TX, R1, C>::SIZE * sizeof(C));
    if (smem > 48 * 1024) {
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, -1, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, +1, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, -1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, +1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, -1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, +1, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, -1, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_pass<T, N, E, TX, +1, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
  }

  // ---- 3D transform legs -------------------------------------------------------------------
  // One GPU:  compact [nf][Kz][Ky][Kxp] <-> [nf][nz][Ky][Kxp] <-> [nf][nz][ny][Kxp].
  // Slabs:    the [nz][Kyl][Kxp] side of the z passes and the ky side of the y passes live in the blocked exchange
  //           layout [peer][nf][nzl][Kyl][Kxp]; the all-to-all in between swaps "all z, my ky" for "my z, all ky".
  void z_inverse(const C* in, long long in_field, C* out, int nf) {
    PassArgs<T> a;
    a.in = in; a.out = out; a.tw = twz;
    a.in_row = a.out_row = Kyl * Kxp;
    a.in_outer = a.out_outer = 0;
    a.in_field = in_field; a.out_field = (long long)nzl * Kyl * Kxp;   // nzl == nz on one GPU; in-block field stride otherwise
    a.inner = Kyl * Kxp; a.lo = bz.lo; a.hi0 = bz.hi0; a.shift = bz.hi0 - bz.lo;
    a.blk_rows = 0; a.blk_stride = 0; a.blk_magic = 0; a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
    if (P_ > 1) { a.blk_rows = nzl; a.blk_stride = (int)blk(nf); a.blk_magic = (unsigned)((0x100000000ULL + nzl - 1) / nzl); }
    blk_out = true;
    prof_begin(KC_ZINV);
    launch_pass<+1, true>(nz, a, 1, nf);
    prof_end();
  }
  void y_inverse(const C* in, C* out, int nf) {
    PassArgs<T> a;
    a.in = in; a.out = out; a.tw = twy;
    a.in_row = a.out_row = Kxp;
    a.in_outer = (long long)Kyl * Kxp; a.out_outer = (long long)ny * Kxp;
    a.in_field = (long long)nzl * Kyl * Kxp; a.out_field = (long long)nzl * ny * Kxp;
    a.inner = Kxp; a.lo = by.lo; a.hi0 = by.hi0; a.shift = by.hi0 - by.lo;
    a.blk_rows = 0; a.blk_stride = 0; a.blk_magic = 0; a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
    if (P_ > 1) { a.blk_rows = Kyl; a.blk_stride = (int)blk(nf); a.blk_magic = (unsigned)((0x100000000ULL + Kyl - 1) / Kyl); }
    blk_out = false;
    prof_begin(KC_YINV);
    launch_pass<+1>(ny, a, nzl, nf);
    prof_end();
  }
  void y_forward(const C* in, C* out, int nf) {
    PassArgs<T> a;
    a.in = in; a.out = out; a.tw = twy;
    a.in_row = a.out_row = Kxp;
    a.in_outer = (long long)ny * Kxp; a.out_outer = (long long)Kyl * Kxp;
    a.in_field = (long long)nzl * ny * Kxp; a.out_field = (long long)nzl * Kyl * Kxp;
    a.inner = Kxp; a.lo = by.lo; a.hi0 = by.hi0; a.shift = by.hi0 - by.lo;
    a.blk_rows = 0; a.blk_stride = 0; a.blk_magic = 0; a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
    if (P_ > 1) { a.blk_rows = Kyl; a.blk_stride = (int)blk(nf); a.blk_magic = (unsigned)((0x100000000ULL + Kyl - 1) / Kyl); }
    blk_out = true;
    prof_begin(KC_YFWD);
    launch_pass<-1>(ny, a, nzl, nf);
    prof_end();
  }
  void z_forward(const C* in, C* out, long long out_field, int nf) {
    PassArgs<T> a;
    a.in = in; a.out = out; a.tw = twz;
    a.in_row = a.out_row = Kyl * Kxp;
    a.in_outer = a.out_outer = 0;
    a.in_field = (long long)nzl * Kyl * Kxp;
    a.out_field = out_field;
    a.inner = Kyl * Kxp; a.lo = bz.lo; a.hi0 = bz.hi0; a.shift = bz.hi0 - bz.lo;
    a.blk_rows = 0; a.blk_stride = 0; a.blk_magic = 0; a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
    if (P_ > 1) { a.blk_rows = nzl; a.blk_stride = (int)blk(nf); a.blk_magic = (unsigned)((0x100000000ULL + nzl - 1) / nzl); }
    blk_out = false;
    prof_begin(KC_ZFWD);
    launch_pass<-1, true>(nz, a, 1, nf);
    prof_end();
  }
  // all-to-all of the blocked layout: piece q (blk(nf) elements) goes to / comes from rank q; the own piece is a local
  // copy.  Two transports: (a) after mhdf_ipc_import, copy-engine pushes straight into the peers' receive buffer over
  // NVLink (no SMs, overlaps the axis passes), closed by a tiny all-reduce as the cross-rank barrier; (b) NCCL
  // send/recv.  `recv` must be R or Q (+ offset).
  // slot >= 0 (pipelined path, peer memory): no collective at all -- every push is followed by a flag store into the
  // receiver's flag array (returns the epoch the receiver has to wait for with wait_flags); the buffers of that path are
  // never aliased, so "the peer is done with its receive buffer" follows from the data dependencies of the step itself
  // (see rhs_pipe).  slot < 0: two cross-rank barriers (before the first push of a leg, after the last push).
  // `len` (default: the whole piece) = elements copied per piece, for pushing a field sub-range of the pieces: send / recv
  // then point at the first field of the range inside piece 0.
  unsigned exchange(const C* send, C* recv, int nf, bool first_of_leg = true, size_t block_elems = 0, int slot = -1, size_t len = 0) {
    const size_t B = block_elems ? block_elems : blk(nf);
    const size_t L = len ? len : B;
    const ncclDataType_t dt = sizeof(T) == 4 ? ncclFloat32 : ncclFloat64;
    const bool flagged = ipc_on && use_flags && slot >= 0;
    unsigned ep = 0;
    prof_begin(KC_EXCH, sc);
    // pushes land in the peers' buffer without the peer posting a receive: before the first push of a leg every rank
    // must be past its last use of that buffer (stream order on each rank + this barrier)
    if (ipc_on && first_of_leg && !flagged) NK(g_nccl.AllReduce(bar_d, bar_d, 1, ncclFloat32, ncclSum, comm, sc));
    CK(cudaMemcpyAsync(recv + (size_t)rank_ * B, send + (size_t)rank_ * B, L * sizeof(C), cudaMemcpyDeviceToDevice, sc));
    if (ipc_on) {
      const bool inR = (recv >= R && recv < R + szR);
      const size_t off = inR ? (size_t)(recv - R) : (size_t)(recv - Q);
      if (flagged) ep = ++epoch_[slot];
      for (int i = 0; i < NCS && i < P_ - 1; ++i) order(sc, cs[i]);
      int k = 0;
      for (int d = 1; d < P_; ++d, ++k) {
        const int q = (rank_ + d) % P_;   // stagger the targets so the pushes of all ranks spread over the links
        C* dst = (inR ? peerR[q] : peerQ[q]) + off + (size_t)rank_ * B;
        CK(cudaMemcpyAsync(dst, send + (size_t)q * B, L * sizeof(C), cudaMemcpyDeviceToDevice, cs[k % NCS]));
        if (flagged) flag_signal(peerF[q] + (size_t)slot * P_ + rank_, ep, cs[k % NCS]);
      }
      for (int i = 0; i < NCS && i < P_ - 1; ++i) order(cs[i], sc);
      if (!flagged) NK(g_nccl.AllReduce(bar_d, bar_d, 1, ncclFloat32, ncclSum, comm, sc));   // every rank's pushes have landed
    } else {
      NK(g_nccl.GroupStart());
      for (int q = 0; q < P_; ++q) {
        if (q == rank_) continue;
        NK(g_nccl.Send(send + (size_t)q * B, 2 * L, dt, q, comm, sc));
        NK(g_nccl.Recv(recv + (size_t)q * B, 2 * L, dt, q, comm, sc));
      }
      NK(g_nccl.GroupEnd());
    }
    prof_end(sc);
    return ep;
  }
  void flag_signal(unsigned* peer_flag, unsigned v, cudaStream_t s) {
#ifdef MHDF_CPU_EMU   // streams are synchronous and kernel launches take turns process-wide: store from the rank's host thread
    (void)s;
    st_flag(peer_flag, v);
#else
    k_flag_set<<<1, 32, 0, s>>>(peer_flag, v);
    ++launches;
#endif
  }
  // the compute stream waits until every peer's piece of (slot, epoch) has landed
  void wait_flags(int slot, unsigned ep) {
    if (ep == 0) return;
#ifdef MHDF_CPU_EMU
    for (int q = 0; q < P_; ++q) if (q != rank_) while ((int)(ld_flag(flags_d + (size_t)slot * P_ + q) - ep) < 0) std::this_thread::yield();
#else
    k_flag_wait<<<1, 32, 0, st>>>(flags_d + (size_t)slot * P_, ep, P_, rank_, ferr_h);
    ++launches;
#endif
  }
  // cross-rank barrier of the compute streams: separates API-level transforms (barrier-mode exchanges on aliased buffers)
  // from the flag-mode pipeline of the time step
  void rank_barrier() {
    if (P_ == 1 || !ipc_on) return;
    order(st, sc);
    NK(g_nccl.AllReduce(bar_d, bar_d, 1, ncclFloat32, ncclSum, comm, sc));
    order(sc, st);
  }
  void check_flags() {
    if (ferr_h && *ferr_h) {
      const int q = *ferr_h - 1;
      *ferr_h = 0;
      throw Err{MHDF_ERR_NCCL, "slab exchange: timed out waiting for the piece of rank " + std::to_string(q)};
    }
  }
  void ipc_export(void* blob) override {
    if (P_ == 1) throw Err{MHDF_ERR_STATE, "peer exchange needs nranks > 1"};
    IpcBlob b;
    std::memset(&b, 0, sizeof b);
    CK(cudaSetDevice(cfg.device));
    CK(cudaIpcGetMemHandle(&b.r, R));
    CK(cudaIpcGetMemHandle(&b.q, Q));
    CK(cudaIpcGetMemHandle(&b.f, flags_d));
    b.device = cfg.device;
    std::memcpy(blob, &b, sizeof b);
  }
  void ipc_import(const void* blobs) override {
    if (P_ == 1) throw Err{MHDF_ERR_STATE, "peer exchange needs nranks > 1"};
    CK(cudaSetDevice(cfg.device));
    peerR.assign(P_, nullptr); peerQ.assign(P_, nullptr); peerF.assign(P_, nullptr);
    const IpcBlob* b = reinterpret_cast<const IpcBlob*>(blobs);
    for (int q = 0; q < P_; ++q) {
      if (q == rank_) { peerR[q] = R; peerQ[q] = Q; peerF[q] = flags_d; continue; }
      void *pr = nullptr, *pq = nullptr, *pf = nullptr;
      CK(cudaIpcOpenMemHandle(&pr, b[q].r, cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle(&pq, b[q].q, cudaIpcMemLazyEnablePeerAccess));
      CK(cudaIpcOpenMemHandle(&pf, b[q].f, cudaIpcMemLazyEnablePeerAccess));
      peerR[q] = reinterpret_cast<C*>(pr); peerQ[q] = reinterpret_cast<C*>(pq); peerF[q] = reinterpret_cast<unsigned*>(pf);
    }
    // copy streams at the highest priority: the one-thread flag kernels behind the pushes must not queue behind the blocks of a
    // big compute kernel
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < NCS; ++i) if (!cs[i]) CK(cudaStreamCreateWithPriority(&cs[i], cudaStreamNonBlocking, prio_hi));
    if (!bar_d) bar_d = dalloc<float>(1);
    CK(cudaStreamSynchronize(st));
    ipc_on = true;
  }
  // fields per exchange chunk: the transposes of one chunk overlap the passes of the next (MHDF_EXCH_CHUNK=0: whole batch)
  int CHUNK = [] { const char* e = getenv("MHDF_EXCH_CHUNK"); return e ? atoi(e) : 3; }();
  int chunk_fields(int nf) const { return (CHUNK > 0 && nf % CHUNK == 0) ? CHUNK : (CHUNK == 0 ? nf : 1); }
  // spectral compact (src, field stride cf) -> x-pass layout in Q.  Uses P (and R when exchanging).
  void to_xlayout(const C* src, int nf) {
    if (P_ == 1) { z_inverse(src, cf, P, nf); y_inverse(P, Q, nf); return; }
    const int fc = chunk_fields(nf), nc = nf / fc;
    const size_t cb = (size_t)P_ * blk(fc);
    std::vector<cudaEvent_t> done(nc);
    for (int c = 0; c < nc; ++c) {
      z_inverse(src + (size_t)c * fc * cf, cf, P + c * cb, fc);
      order(st, sc);
      exchange(P + c * cb, R + c * cb, fc, c == 0);
      done[c] = dep_ev[dep_next++ % dep_ev.size()];
      CK(cudaEventRecord(done[c], sc));
    }
    for (int c = 0; c < nc; ++c) {
      CK(cudaStreamWaitEvent(st, done[c], 0));
      y_inverse(R + c * cb, Q + (size_t)c * fc * nzl * ny * Kxp, fc);
    }
  }
  // x-pass layout in `src` (R or Q) -> compact spectral in dst (field stride cf).  Uses P and, when exchanging, `via`.
  void from_xlayout(const C* src, C* via, C* dst, int nf) {
    if (P_ == 1) { y_forward(src, P, nf); z_forward(P, dst, cf, nf); return; }
    const int fc = chunk_fields(nf), nc = nf / fc;
    const size_t cb = (size_t)P_ * blk(fc);
    std::vector<cudaEvent_t> done(nc);
    for (int c = 0; c < nc; ++c) {
      y_forward(src + (size_t)c * fc * nzl * ny * Kxp, P + c * cb, fc);
      order(st, sc);
      exchange(P + c * cb, via + c * cb, fc, c == 0);
      done[c] = dep_ev[dep_next++ % dep_ev.size()];
      CK(cudaEventRecord(done[c], sc));
    }
    // dst may alias src (R): every forward-y pass has been issued before the first forward-z pass writes
    for (int c = 0; c < nc; ++c) {
      CK(cudaStreamWaitEvent(st, done[c], 0));
      z_forward(via + c * cb, dst + (size_t)c * fc * cf, cf, fc);
    }
  }
  // kr = 0 plane of the stage input from every rank (the symmetrised diffusion operand needs the mirror mode).
  // Issued on the communication stream; `mirror_ready` is waited for right before the consumer kernel.
  cudaEvent_t mirror_ready = nullptr;
  void gather_mirror(const C* S) {
    if (P_ == 1) return;
    const long long n = (long long)F * Kz * Kyl;
    MHDF_LAUNCH((k_plane<T>), (int)((n + 255) / 256), 256, 0, st, geom(), S, plane_loc);
    ++launches;
    CK(cudaGetLastError());
    order(st, sc);
    NK(g_nccl.AllGather(plane_loc, plane_all, 2 * (size_t)n, sizeof(T) == 4 ? ncclFloat32 : ncclFloat64, comm, sc));
    mirror_ready = dep_ev[dep_next++ % dep_ev.size()];
    CK(cudaEventRecord(mirror_ready, sc));
  }
  void wait_mirror() {
    if (P_ > 1 && mirror_ready) { CK(cudaStreamWaitEvent(st, mirror_ready, 0)); mirror_ready = nullptr; }
  }
  // global sums / maxima of the x-kernel reductions, then the host copy
  cudaEvent_t red_ev = nullptr, red_own = nullptr;   // red_own: dedicated event (it is waited for a whole step after its record)
  void finish_red() {
    if (P_ == 1) { CK(cudaMemcpyAsync(red_h, red_d, sizeof(XRed), cudaMemcpyDeviceToHost, st)); return; }
    order(st, sc);
    NK(g_nccl.AllReduce(red_d->sumsq, red_d->sumsq, 8, ncclFloat64, ncclSum, comm, sc));
    NK(g_nccl.AllReduce(red_d->maxsq, red_d->maxsq, 6, ncclUint64, ncclMax, comm, sc));
    CK(cudaMemcpyAsync(red_h, red_d, sizeof(XRed), cudaMemcpyDeviceToHost, sc));
    CK(cudaEventRecord(red_own, sc));
    red_ev = red_own;    // the next reset of red_d must not overtake the copy (waited for in red_reset, not here: the compute
                         // stream must not stall behind the pushes queued on the communication stream)
  }
  void red_reset() {
    if (red_ev) { CK(cudaStreamWaitEvent(st, red_ev, 0)); red_ev = nullptr; }
    CK(cudaMemsetAsync(red_d, 0, sizeof(XRed), st));
  }

  XArgs<T> xargs() const {
    XArgs<T> a;
    a.in = Q; a.out = R; a.tw = twx; a.real_io = nullptr;
    a.in_field = a.out_field = (long long)nzl * ny * Kxp;
    a.real_field = (long long)nx * ny * nzl;
    a.rows = (long long)ny * nzl;
    a.Kx = Kx; a.Kxp = Kxp;
    a.scale = (T)(1.0 / ((double)nx * ny * nz));
    a.red = nullptr;
    a.vp = nullptr; a.vp_field = 0; a.vp_eta = (T)1;
    a.kxv = kxv;
    return a;
  }
  // penalised RHS evaluations: the x kernel reads chi, U0 (B0) next to the row set; eta = clock.dt * 13/7 (VPSolver.jl:23)
  void set_vp(XArgs<T>& xa, size_t real_off) const {
    if (nd_on) {   // the negative-damping profiles f_i take the place of the penalisation fields
      xa.vp = (nd_P != 0) ? nd_d + real_off : nullptr;
      xa.vp_field = (long long)nx * ny * nzl;
      return;
    }
    if (!vp_on) return;
    xa.vp = vp_d + real_off;
    xa.vp_field = (long long)nx * ny * nzl;
    xa.vp_eta = dt_ * (T)13 / (T)7;
  }

  // ---- z-chunk pipelined slab path (opt-in: MHDF_ZCHUNKS = 2, 4, ...) ---------------------------------------------
  // The local z slab is cut into NZC chunks; exchange pieces are [chunk][peer][field][z''][ky'][kx].  After the inverse z
  // pass the inverse pushes of all chunks are queued on the communication stream; as soon as chunk c has arrived, its
  // inverse y pass, fused x pass and forward y pass run and its forward pushes are queued -- so the x pass of one chunk
  // overlaps the pushes of the others.  Only the first inverse and the last forward exchange stay exposed.
  // Separate buffers keep pushes from peers (which land in R / Q unannounced) away from live data:
  //   P inverse send, R inverse receive, Xin x-pass input (later the product spectra), Xout x-pass output,
  //   P2 forward send, Q forward receive.
  // NOT YET VERIFIED ON HARDWARE (written after the round's GPU budget was spent); off unless MHDF_ZCHUNKS is set.
  // EMHD x kernel: the shared-memory-multiplier form (k_xfused_emhd2) is the default -- 32.6 ms against 45.3 ms per 512^3 step
  // for the register form on B200 (profiles/README.md); MHDF_EMHD2=0 selects the register form (bit-identical results)
  bool emhd2 = [] { const char* e = getenv("MHDF_EMHD2"); return !e || atoi(e) != 0; }();
  // z chunks of the pipelined slab path: MHDF_ZCHUNKS, or (unset) chosen from the size of the pushes -- 4 chunks when one
  // chunk's inverse pushes still move >= 64 MB per round, else 2 (measured on B200s: 8 chunks lose to 4 at 512^3 and 1024^3,
  // copy-engine pushes below ~30 MB run at less than half of the NVLink rate; profiles/r02_c3_*)
  int zchunks = [] { const char* e = getenv("MHDF_ZCHUNKS"); const int n = e ? atoi(e) : 0; return n < 0 ? 0 : n; }();
  void choose_chunks() {
    if (P_ == 1) { zchunks = 1; return; }
    if (zchunks > 0) return;
    auto round_bytes = [&](int nzc) { return (double)nin * (nzl / nzc) * Kyl * Kxp * sizeof(C) * (P_ - 1); };
    zchunks = (nzl % 4 == 0 && nzl / 4 >= 2 && round_bytes(4) >= 64e6) ? 4 : 2;
  }
  C *Xin = nullptr, *Xout = nullptr, *P2 = nullptr;
  bool pipe_ok() const {
    if (P_ == 1 || zchunks <= 1 || zchunks > 8 || nzl % zchunks != 0) return false;
    const int zc = nzl / zchunks, rb = 1024 / nx > 1 ? 1024 / nx : 1;
    if (zc < 2) return false;   // one plane per chunk: the second-level divisor would be 1 (see init)
    if ((long long)(nin > nout ? nin : nout) * nz * Kyl * Kxp >= (1LL << 31)) return false;   // 32-bit row offsets
    return ((long long)ny * zc) % rb == 0;
  }
  void pipe_alloc() {
    if (Xin) return;
    const size_t e_zy = (size_t)nzl * ny * Kxp, e_zK = (size_t)nz * Kyl * Kxp;
    Xin = dalloc<C>((size_t)(nin > nout ? nin : nout) * e_zy);
    Xout = dalloc<C>((size_t)nout * e_zy);
    P2 = dalloc<C>((size_t)nout * e_zK);
  }
  void set_blk(PassArgs<T>& a, int rows, size_t stride) {
    a.blk_rows = rows; a.blk_stride = (int)stride; a.blk_magic = (unsigned)((0x100000000ULL + rows - 1) / rows);
  }
  // field groups of the pipelined exchanges: the pieces of one (chunk, peer) are pushed group by group so that the z passes
  // at both ends of an evaluation overlap the first / last pushes (MHDF_FGROUPS=0: one group)
  // Field groups of the pipelined exchanges: the pieces of a (chunk, peer) can be pushed group by group so that the z passes at
  // both ends of an evaluation overlap the first / last pushes.  MHDF_FGROUPS: 0 never, 1 every chunk, 2 only where it shortens
  // the critical path (the first inverse chunk and the last forward chunk); unset: 2 when a group's piece is still >= 16 MB,
  // and every chunk when it is >= 100 MB (measured on B200s: copy-engine pushes of ~20 MB run at less than half of the rate of
  // 100 MB pushes, so small pieces cost more than the overlap returns; profiles/r02_c3_*, r02_c4_*)
  int fgroups_env = [] { const char* e = getenv("MHDF_FGROUPS"); return e ? atoi(e) : -1; }();
  int group_mode(int nf, int g) const {
    if (fgroups_env >= 0) return fgroups_env;
    const double piece = (double)(nf / g) * (nzl / zchunks) * Kyl * Kxp * sizeof(C);
    return piece >= 100e6 ? 1 : (piece >= 16e6 ? 2 : 0);
  }
  int groups_max(int nf) const {
    for (int g = 4; g > 1; --g) if (nf % g == 0 && nf / g >= 3) return g;
    return 1;
  }
  // pushes of one exchange without joining the copy streams: every peer's stream runs its pushes back to back, ordered only
  // behind the kernel that produced the data (`ready`) -- no per-round rendezvous of the seven streams (MHDF_NOJOIN=0: joined)
  bool nojoin = [] { const char* e = getenv("MHDF_NOJOIN"); return !e || atoi(e) != 0; }();
  cudaEvent_t mark(cudaStream_t s) {
    cudaEvent_t e = dep_ev[dep_next++ % dep_ev.size()];
    CK(cudaEventRecord(e, s));
    return e;
  }
  // One push round of the pipelined path: elements [0, L) of every (peer) piece (pieces B apart) from `send` to the same place in
  // the peers' receive buffer `recv`; `ready` = the producing kernel has finished.  Returns the epoch to wait for.
  unsigned push_round(const C* send, C* recv, size_t B, size_t L, int slot, cudaEvent_t ready) {
    if (!(ipc_on && use_flags && nojoin)) {
      CK(cudaStreamWaitEvent(sc, ready, 0));
      return exchange(send, recv, 0, false, B, slot, L);
    }
    const unsigned ep = ++epoch_[slot];
    const bool inR = (recv >= R && recv < R + szR);
    const size_t off = inR ? (size_t)(recv - R) : (size_t)(recv - Q);
    CK(cudaStreamWaitEvent(sc, ready, 0));
    CK(cudaMemcpyAsync(recv + (size_t)rank_ * B, send + (size_t)rank_ * B, L * sizeof(C), cudaMemcpyDeviceToDevice, sc));
    int k = 0;
    for (int d = 1; d < P_; ++d, ++k) {
      const int q = (rank_ + d) % P_;
      cudaStream_t s = cs[k % NCS];
      CK(cudaStreamWaitEvent(s, ready, 0));
      if (k == 0) prof_begin(KC_EXCH, s);
      C* dst = (inR ? peerR[q] : peerQ[q]) + off + (size_t)rank_ * B;
      CK(cudaMemcpyAsync(dst, send + (size_t)q * B, L * sizeof(C), cudaMemcpyDeviceToDevice, s));
      flag_signal(peerF[q] + (size_t)slot * P_ + rank_, ep, s);
      if (k == 0) prof_end(s);
    }
    pushes_pending = true;
    return ep;
  }
  bool pushes_pending = false;
  // my own pushes out of the send buffers must have completed before the next evaluation overwrites those buffers
  void drain_pushes() {
    if (!pushes_pending) return;
    for (int i = 0; i < NCS && i < P_ - 1; ++i) CK(cudaStreamWaitEvent(st, mark(cs[i]), 0));
    pushes_pending = false;
  }
  void rhs_pipe(const C* Sin, SpecArgs<T> sa, bool want_red) {
    pipe_alloc();
    want_red = want_red || (nd_on && nd_P != 0);   // the negative-damping force is normalised by a reduction of every evaluation
    const int NZC = zchunks, zc = nzl / NZC;
    const size_t Bi = (size_t)nin * zc * Kyl * Kxp, Bo = (size_t)nout * zc * Kyl * Kxp;   // one (chunk, peer) piece
    const long long fld = (long long)zc * Kyl * Kxp;                                      // field stride inside a piece
    const int Gim = groups_max(nin), Gom = groups_max(nout);
    const int mi = Gim > 1 ? group_mode(nin, Gim) : 0, mo = Gom > 1 ? group_mode(nout, Gom) : 0;
    auto Gi_of = [&](int c) { return (mi == 1 || (mi == 2 && c == 0)) ? Gim : 1; };          // groups of inverse chunk c
    auto Go_of = [&](int c) { return (mo == 1 || (mo == 2 && c == NZC - 1)) ? Gom : 1; };    // groups of forward chunk c
    const int Gz = mi ? Gim : 1, fgz = nin / Gz;      // the inverse z pass runs group by group whenever any chunk is grouped
    const int Gf = mo ? Gom : 1, fgf = nout / Gf;     // ... and so does the forward z pass
    const C* zin = Sin;
    if (phys != MHDF_EMHD) gather_mirror(Sin);
    if (phys == MHDF_EMHD) {
      prof_begin(KC_DERIVE);
      MHDF_LAUNCH((k_emhd_derive<T>), spec_grid(), 256, 0, st, geom(), Sin, D);
      ++launches;
      CK(cudaGetLastError());
      prof_end();
      zin = D;
    }
    sa.g = geom();
    sa.Sin = Sin;
    sa.nu = (T)cfg.nu; sa.eta = (T)cfg.eta; sa.n_nu = cfg.n_nu;
    sa.force = fmask ? force : nullptr; sa.fmask = fmask;
    next_a99(sa);
    if (want_red) red_reset();
    drain_pushes();
    // Flag mode (peer memory): no collective in this function.  Why a peer's receive buffer is free when my push arrives:
    //  R (inverse receive) of peer q is read by its inverse y passes; my next inverse push follows my spectral update, which
    //    waited for q's forward pieces of ALL chunks, each sent after q's forward y pass of that chunk, i.e. after q read R;
    //  Q (forward receive) of peer q is read by its forward z pass; my next forward push of chunk c follows my inverse y pass
    //    of chunk c, which waited for q's inverse piece of the next evaluation, sent after q's spectral update, i.e. after
    //    q's forward z pass read Q.
    std::vector<cudaEvent_t> zdone(Gz);
    for (int g = 0; g < Gz; ++g) {   // inverse z pass, one field group at a time, into the two-level send layout
      PassArgs<T> a;
      a.in = zin + (size_t)g * fgz * cf; a.out = P + (size_t)g * fgz * fld; a.tw = twz;
      a.in_row = a.out_row = Kyl * Kxp;
      a.in_outer = a.out_outer = 0;
      a.in_field = cf; a.out_field = fld;
      a.inner = Kyl * Kxp; a.lo = bz.lo; a.hi0 = bz.hi0; a.shift = bz.hi0 - bz.lo;
      set_blk(a, nzl, Bi);
      a.blk2_rows = zc; a.blk2_stride = (int)((size_t)P_ * Bi); a.blk2_magic = (unsigned)((0x100000000ULL + zc - 1) / zc);
      blk_out = true;
      prof_begin(KC_ZINV);
      launch_pass<+1, true>(nz, a, 1, fgz);
      prof_end();
      zdone[g] = mark(st);
    }
    struct Round { int slot; unsigned ep; };
    std::vector<std::vector<Round>> inv_r(NZC), fwd_r(NZC);
    std::vector<cudaEvent_t> inv_local(NZC);
    for (int c = 0; c < NZC; ++c) {   // pushes in the order the consumers need them: chunk by chunk, group by group
      const int G = Gi_of(c), fg = nin / G;
      for (int g = 0; g < G; ++g) {
        const size_t o = (size_t)c * P_ * Bi + (size_t)g * fg * fld;
        // data of group g of this chunk is complete once the z passes of the field groups it spans are done
        const cudaEvent_t ready = zdone[G == 1 ? Gz - 1 : g];
        inv_r[c].push_back({c * 4 + g, push_round(P + o, R + o, Bi, (size_t)fg * fld, c * 4 + g, ready)});
      }
      inv_local[c] = mark(sc);
    }
    std::vector<cudaEvent_t> fwd_local;
    for (int c = 0; c < NZC; ++c) {
      const size_t zoff = (size_t)c * zc * ny * Kxp;
      CK(cudaStreamWaitEvent(st, inv_local[c], 0));
      for (const Round& r : inv_r[c]) wait_flags(r.slot, r.ep);
      {   // inverse y pass of chunk c: received pieces -> x-pass layout
        PassArgs<T> a;
        a.in = R + (size_t)c * P_ * Bi; a.out = Xin + zoff; a.tw = twy;
        a.in_row = a.out_row = Kxp;
        a.in_outer = (long long)Kyl * Kxp; a.out_outer = (long long)ny * Kxp;
        a.in_field = fld; a.out_field = (long long)nzl * ny * Kxp;
        a.inner = Kxp; a.lo = by.lo; a.hi0 = by.hi0; a.shift = by.hi0 - by.lo;
        set_blk(a, Kyl, Bi);
        a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
        blk_out = false;
        prof_begin(KC_YINV);
        launch_pass<+1>(ny, a, zc, nin);
        prof_end();
      }
      XArgs<T> xa = xargs();
      xa.in = Xin + zoff; xa.out = Xout + zoff;
      xa.real_io = bst ? bst + (size_t)c * zc * ny * nx : nullptr;
      set_vp(xa, (size_t)c * zc * ny * nx);
      xa.rows = (long long)ny * zc;
      xa.red = want_red ? red_d : nullptr;
      prof_begin(KC_XFUSED);
      launch_xfused(xa);
      prof_end();
      if (want_red && c == NZC - 1) finish_red();
      const int G = Go_of(c), fg = nout / G;
      for (int g = 0; g < G; ++g) {   // forward y pass of chunk c (group by group where grouped), each followed by its push
        PassArgs<T> a;
        a.in = Xout + zoff + (size_t)g * fg * nzl * ny * Kxp; a.out = P2 + (size_t)c * P_ * Bo + (size_t)g * fg * fld; a.tw = twy;
        a.in_row = a.out_row = Kxp;
        a.in_outer = (long long)ny * Kxp; a.out_outer = (long long)Kyl * Kxp;
        a.in_field = (long long)nzl * ny * Kxp; a.out_field = fld;
        a.inner = Kxp; a.lo = by.lo; a.hi0 = by.hi0; a.shift = by.hi0 - by.lo;
        set_blk(a, Kyl, Bo);
        a.blk2_rows = 0; a.blk2_stride = 0; a.blk2_magic = 0;
        blk_out = true;
        prof_begin(KC_YFWD);
        launch_pass<-1>(ny, a, zc, fg);
        prof_end();
        const size_t o = (size_t)c * P_ * Bo + (size_t)g * fg * fld;
        fwd_r[c].push_back({32 + c * 4 + g, push_round(P2 + o, Q + o, Bo, (size_t)fg * fld, 32 + c * 4 + g, mark(st))});
        fwd_local.push_back(mark(sc));
      }
    }
    for (cudaEvent_t e : fwd_local) CK(cudaStreamWaitEvent(st, e, 0));   // own pieces (local copies) are in place
    for (int g = 0; g < Gf; ++g) {   // forward z pass of a field group as soon as its pieces of every chunk have landed
      for (int c = 0; c < NZC; ++c) {
        if ((int)fwd_r[c].size() == 1) { if (g == 0) wait_flags(fwd_r[c][0].slot, fwd_r[c][0].ep); }   // whole-chunk push: wait once
        else wait_flags(fwd_r[c][g].slot, fwd_r[c][g].ep);
      }
      PassArgs<T> a;
      a.in = Q + (size_t)g * fgf * fld; a.out = Xin + (size_t)g * fgf * cf; a.tw = twz;   // compact product spectra go to Xin (free by now)
      a.in_row = a.out_row = Kyl * Kxp;
      a.in_outer = a.out_outer = 0;
      a.in_field = fld; a.out_field = cf;
      a.inner = Kyl * Kxp; a.lo = bz.lo; a.hi0 = bz.hi0; a.shift = bz.hi0 - bz.lo;
      set_blk(a, nzl, Bo);
      a.blk2_rows = zc; a.blk2_stride = (int)((size_t)P_ * Bo); a.blk2_magic = (unsigned)((0x100000000ULL + zc - 1) / zc);
      blk_out = false;
      prof_begin(KC_ZFWD);
      launch_pass<-1, true>(nz, a, 1, fgf);
      prof_end();
    }
    sa.P = Xin;
    wait_mirror();
    launch_spectral(sa);
  }

  // calcF! = A99ForceDriving!: every RHS evaluation draws fresh random numbers (counter word = evaluation number)
  void next_a99(SpecArgs<T>& sa) {
    sa.a99 = a99;
    if (a99.variant == A99_OFF) return;
    sa.a99.call_lo = (unsigned)a99_call; sa.a99.call_hi = (unsigned)(a99_call >> 32);
    ++a99_call;
  }
  // One RHS evaluation of stage input Sin, finished by the spectral kernel in mode sa.mode.
  void rhs(const C* Sin, SpecArgs<T> sa, bool want_red) {
    forcing_callback(Sin);
    if (pipe_ok()) { rhs_pipe(Sin, sa, want_red); return; }
    want_red = want_red || (nd_on && nd_P != 0);
    const C* zin = Sin;
    if (phys != MHDF_EMHD) gather_mirror(Sin);
    if (phys == MHDF_EMHD) {
      prof_begin(KC_DERIVE);
      MHDF_LAUNCH((k_emhd_derive<T>), spec_grid(), 256, 0, st, geom(), Sin, D);
      ++launches;
      CK(cudaGetLastError());
      prof_end();
      zin = D;
    }
    sa.g = geom();
    sa.Sin = Sin;
    sa.nu = (T)cfg.nu; sa.eta = (T)cfg.eta; sa.n_nu = cfg.n_nu;
    sa.force = fmask ? force : nullptr; sa.fmask = fmask;
    next_a99(sa);
    if (want_red) red_reset();
    to_xlayout(zin, nin);
    XArgs<T> xa = xargs();
    xa.real_io = bst;
    set_vp(xa, 0);
    if (want_red) xa.red = red_d;
    prof_begin(KC_XFUSED);
    launch_xfused(xa);
    prof_end();
    if (want_red) finish_red();
    C* spec = (P_ > 1) ? R : Q;           // forward-z output (compact product spectra)
    from_xlayout(R, Q, spec, nout);
    sa.P = spec;
    wait_mirror();
    launch_spectral(sa);
  }
  template <int PHYS, bool A99, int VP> void launch_spectral_m(const SpecArgs<T>& sa) {
    const unsigned plane = (unsigned)Kxp * (unsigned)Kyl;
    const dim3 grid((plane + 255u) / 256u, (unsigned)Kz);
    switch (sa.mode) {
      case STEP_CALCN: MHDF_LAUNCH((k_spectral<T, PHYS, STEP_CALCN, A99, VP>), grid, 256, 0, st, sa); break;
      case STEP_RK4_1: MHDF_LAUNCH((k_spectral<T, PHYS, STEP_RK4_1, A99, VP>), grid, 256, 0, st, sa); break;
      case STEP_RK4_2: MHDF_LAUNCH((k_spectral<T, PHYS, STEP_RK4_2, A99, VP>), grid, 256, 0, st, sa); break;
      case STEP_RK4_3: MHDF_LAUNCH((k_spectral<T, PHYS, STEP_RK4_3, A99, VP>), grid, 256, 0, st, sa); break;
      case STEP_RK4_4: MHDF_LAUNCH((k_spectral<T, PHYS, STEP_RK4_4, A99, VP>), grid, 256, 0, st, sa); break;
      default:         MHDF_LAUNCH((k_spectral<T, PHYS, STEP_LSRK, A99, VP>), grid, 256, 0, st, sa); break;
    }
  }
  void launch_spectral(SpecArgs<T>& sa) {
    prof_begin(KC_SPEC);
    const bool driven = (phys == MHDF_MHD) && sa.a99.variant != A99_OFF;   // A99ForceDriving! acts on the MHD path only
    if (nd_on && nd_P != 0) {   // NDForceDriving!: the normalisation sum of this evaluation is complete (and all-reduced) by now
      if (P_ > 1 && red_ev) CK(cudaStreamWaitEvent(st, red_ev, 0));
      sa.nd_sum = &red_d->nd;
      sa.nd_P = nd_P / ((cfg.Lx / nx) * (cfg.Ly / ny) * (cfg.Lz / nz));
      launch_spectral_m<PHYS_MHD, false, 2>(sa);
    } else if (vp_on) {   // penalised runs: the product buffer carries the penalisation spectra as well
      if (driven) launch_spectral_m<PHYS_MHD, true, true>(sa);
      else if (phys == MHDF_MHD) launch_spectral_m<PHYS_MHD, false, true>(sa);
      else launch_spectral_m<PHYS_HD, false, true>(sa);
    } else {
      if (driven) launch_spectral_m<PHYS_MHD, true, false>(sa);
      else if (phys == MHDF_MHD) launch_spectral_m<PHYS_MHD, false, false>(sa);
      else if (phys == MHDF_HD) launch_spectral_m<PHYS_HD, false, false>(sa);
      else launch_spectral_m<PHYS_EMHD, false, false>(sa);
    }
    ++launches;
    CK(cudaGetLastError());
    prof_end();
  }
  int spec_grid() const {
    long long b = (cf + 255) / 256;
    long long cap = (long long)nsm * 16;
    return (int)(b < cap ? b : cap);
  }

  double red_max(int i) const {   // max f^2 in the problem's precision (XRed::maxsq holds the bit pattern of a T)
    T v;
    std::memcpy(&v, &red_h->maxsq[i], sizeof(T));
    return (double)v;
  }
  void absorb_red() {   // after a stream sync: stale vars statistics of the last RHS evaluation
    for (int i = 0; i < 6; ++i) {
      st_sum[i] = red_h->sumsq[i];
      st_max[i] = red_max(i);
    }
    st_cross = red_h->cross;
  }

  SpecArgs<T> blank_args() const {
    SpecArgs<T> sa;
    std::memset(&sa, 0, sizeof sa);
    sa.dt = dt_;
    return sa;
  }

  void one_step() {
    const T dt = dt_;
    if (cfg.stepper == MHDF_HM89) { hm89_step(); return; }
    if (cfg.stepper == MHDF_RK4) {
      // registers: Y = reg[iY]; two stage buffers and the accumulator are the other three
      int o[3], n = 0;
      for (int i = 0; i < 4; ++i) if (i != iY) o[n++] = i;
      C *Y = reg[iY], *S0 = reg[o[0]], *S1 = reg[o[1]], *A = reg[o[2]];
      SpecArgs<T> sa = blank_args();
      sa.Y = Y; sa.A = A;
      sa.mode = STEP_RK4_1; sa.ca = dt / (T)6; sa.cs = dt / (T)2; sa.Sout = S0;
      t_eval = (double)t_;                          // the stage times FourierFlows hands to calcN! (clock.t and dt are T)
      rhs(Y, sa, false);
      sa.mode = STEP_RK4_2; sa.ca = dt / (T)3; sa.cs = dt / (T)2; sa.Sout = S1;
      t_eval = (double)(T)(t_ + dt / (T)2);
      rhs(S0, sa, false);
      sa.mode = STEP_RK4_3; sa.ca = dt / (T)3; sa.cs = dt; sa.Sout = S0;
      rhs(S1, sa, false);
      t_eval = (double)(T)(t_ + dt);
      sa.mode = STEP_RK4_4; sa.ca = dt / (T)6; sa.cs = 0; sa.Sout = Y;
      rhs(S0, sa, true);
      iStale = o[0];
    } else {
      static const double LA[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                   -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
      static const double LB[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                   1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                   2277821191437.0 / 14882151754819.0};
      static const double LC[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                                   2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
      // registers: sol ping-pongs between two buffers, S2 is the third
      int o[2], n = 0;
      for (int i = 0; i < 3; ++i) if (i != iY) o[n++] = i;
      int cur = iY, oth = o[0];
      C* S2 = reg[o[1]];
      for (int i = 0; i < 5; ++i) {
        SpecArgs<T> sa = blank_args();
        sa.mode = STEP_LSRK; sa.A = S2; sa.ca = (T)LA[i]; sa.cs = (T)LB[i]; sa.first = (i == 0);
        sa.Sout = reg[oth];
        t_eval = (double)(T)(t_ + (T)LC[i] * dt);
        rhs(reg[cur], sa, i == 4);
        int tmp = cur; cur = oth; oth = tmp;
      }
      // after 5 swaps `cur` holds the new sol, `oth` the 5th stage input (= stale vars source)
      iY = cur;
      iStale = oth;
      // keep S2 where it is: the third register
    }
    t_ = t_ + dt;
    step_ += 1;
  }


  // ---- HM89TimeStepper (timestepper/HM89.jl:23-199; Problem(...; EMHD = true, stepper = "HM89"), Problems.jl:124-126) ------------
  // One step = LSRK3 predictor (3 Hall-term evaluations) -> DivFreeCorrection! -> fixed-point iteration of the implicit midpoint
  // rule (one evaluation + one three-field inverse transform + one max reduction per iteration, until max |B^n - B^1| <= 5e-4) ->
  // RK3linearterm! (resistive term, three explicit stages that also read what the loop left in F0 / F1, like the reference) ->
  // DivFreeCorrection! -> vars.b = irfft(sol).  Registers: sol = reg[iY], then F0, F1, B0, B1 (B^n is formed in place).
  // The reference writes B^n - B^1 in real space INTO vars.bx/by/bz (HM89.jl:49,76-78): the next evaluation of the loop reads that
  // difference as its stale b -- reproduced (the inverse transforms land in the stale-b array of the EMHD x kernel).
  // One deviation: the closing ldiv!(vars.b, sol) of the reference sees the aliased band that RK3linearterm! left in sol
  // (c2 dt N on the modes dealias! removes at the next evaluation); the compact state has no such band, vars.b is the dealiased field.
  long long hm_iters = 0;   // fixed-point iterations of the last step
  double hm_eps = 0;        // its last error norm
  static constexpr int HM_MAX_ITERS = 10000;
  void hm_launch(const Hm89Args<T>& a) {
    MHDF_LAUNCH((k_hm89<T>), spec_grid(), 256, 0, st, geom(), a);
    ++launches;
    CK(cudaGetLastError());
  }
  void hm_eval(C* N, const C* S, bool want_red) {   // N = EMHDcalcN!(S) (no forcing, no resistive term: pgen.jl:164-171)
    SpecArgs<T> sa = blank_args();
    sa.mode = STEP_CALCN; sa.Nout = N;
    rhs(S, sa, want_red);
  }
  void hm_stage(C* Fa, const C* G, C* S, double b, double c, bool first, bool lin) {
    Hm89Args<T> a;
    std::memset(&a, 0, sizeof a);
    a.op = HM_STAGE; a.lin = lin ? 1 : 0;
    a.F = Fa; a.G = G; a.S = S;
    a.eta = cfg.eta; a.b = b;
    a.dt = dt_; a.c = (T)c; a.c2 = first ? dt_ : (T)1;
    hm_launch(a);
  }
  void hm_divfree(C* S) {
    MHDF_LAUNCH((k_divclean<T>), spec_grid(), 256, 0, st, geom(), S);
    ++launches;
    CK(cudaGetLastError());
  }
  void hm89_step() {
    C *S = reg[iY];
    C* o[4];
    { int n = 0; for (int i = 0; i < 5; ++i) if (i != iY) o[n++] = reg[i]; }
    C *F0 = o[0], *F1 = o[1], *B0 = o[2], *B1 = o[3];
    const size_t bytes = (size_t)F * cf * sizeof(C);
    const double c1 = 1.0 / 3.0, c2 = 15.0 / 16.0, c3 = 8.0 / 15.0;
    CK(cudaMemcpyAsync(B0, S, bytes, cudaMemcpyDeviceToDevice, st));                    // copyto!(B0, sol)
    hm_eval(F0, S, false); hm_stage(F0, nullptr, S, 0.0, c1, true, false);              // LSRK3substeps!
    hm_eval(F1, S, false); hm_stage(F1, F0, S, 5.0 / 9.0, c2, false, false);
    hm_eval(F0, S, false); hm_stage(F0, F1, S, 153.0 / 128.0, c3, false, false);
    hm_divfree(S);
    CK(cudaMemcpyAsync(B1, S, bytes, cudaMemcpyDeviceToDevice, st));                    // copyto!(B1, sol); dealias!(B1)
    T* re = reinterpret_cast<T*>(R);
    const size_t n = (size_t)nx * ny * nzl;
    double eps = 1.0;
    hm_iters = 0;
    while (eps > 5e-4) {
      Hm89Args<T> a;
      std::memset(&a, 0, sizeof a);
      a.op = HM_HALF; a.S = S; a.B0 = B0; a.B1 = B1;
      hm_launch(a);                                                                    // B_half = (B0 + B1) 0.5
      hm_eval(F1, S, true);                                                            // the Hall term of B_half; its curl B maxima feed getCFL!
      sync_all();
      check_flags();
      absorb_red();
      a.op = HM_FIXED; a.F = F0; a.G = F1; a.dt = dt_;
      hm_launch(a);                                                                    // B^n = B0 + dt N; F0 = B^n - B^1; B^1 = B^n
      for (int i = 0; i < 3; ++i) {                                                    // vars.b <- irfft(B^n - B^1)
        to_xlayout(F0 + (size_t)i * cf, 1);
        XArgs<T> xa = xargs();
        xa.real_io = re; xa.in = Q; xa.out = nullptr;
        launch_xplain<+1>(xa);
        CK(cudaMemcpyAsync(bst + (size_t)i * n, re, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
      }
      red_reset();
      MHDF_LAUNCH((k_norm3_max<T>), spec_grid(), 256, 0, st, bst, bst + n, bst + 2 * n, (long long)n, &red_d->maxsq[0]);
      ++launches;
      CK(cudaGetLastError());
      finish_red();
      sync_all();
      check_flags();
      eps = red_max(0);
      ++hm_iters;
      if (!std::isfinite(eps)) break;                                                  // NaN ends the reference's loop too; the caller's NaN check reports it
      if (hm_iters >= HM_MAX_ITERS) throw Err{MHDF_ERR_STATE, "HM89: the fixed-point iteration did not converge (the reference would loop forever)"};
    }
    hm_eps = eps;
    CK(cudaMemcpyAsync(S, B1, bytes, cudaMemcpyDeviceToDevice, st));                    // copyto!(sol, B1)
    hm_stage(F0, nullptr, S, 0.0, c1, true, true);                                      // RK3linearterm! with calcF! = nothingfunction
    hm_stage(F1, F0, S, 5.0 / 9.0, c2, false, true);
    hm_stage(F0, F1, S, 153.0 / 128.0, c3, false, true);
    hm_divfree(S);
    iStale = iY;
    refresh_vars(0);                                                                   // vars.b = irfft(sol): stale b, maxima, energies
    if (!std::isfinite(eps)) st_sum[3] = eps;                                          // surfaces as MHDF_ERR_NONFINITE in step()
    t_ = t_ + dt_;
    step_ += 1;
  }

  void stepper_stats(long long* iters, double* eps) const override {
    if (iters) *iters = hm_iters;
    if (eps) *eps = hm_eps;
  }
  void step(int n) override {
    CK(cudaSetDevice(cfg.device));
    rank_barrier();
    for (int i = 0; i < n; ++i) one_step();
    sync_all();
    check_flags();
    if (n > 0) {
      if (cfg.stepper != MHDF_HM89) absorb_red();   // HM89 keeps its own statistics (curl B of the last evaluation, b of the closing ldiv!)
      check_finite();
    }
  }
  void check_finite() {
    for (int i = 0; i < 6; ++i)
      if (!std::isfinite(st_sum[i])) throw Err{MHDF_ERR_NONFINITE, "detected NaN! Quit the simulation right now."};
  }
  void step_timed(int n, double* ms) override {
    CK(cudaSetDevice(cfg.device));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    rank_barrier();
    sync_all();
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < n; ++i) one_step();
    CK(cudaEventRecord(b, st));
    sync_all();
    check_flags();
    float f = 0;
    CK(cudaEventElapsedTime(&f, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms = f;
    if (n > 0) { if (cfg.stepper != MHDF_HM89) absorb_red(); check_finite(); }
  }

  void calcN(void* p) override {
    CK(cudaSetDevice(cfg.device));
    // N goes to a register that is dead between steps
    int o = -1;
    const int nreg = nreg_();
    for (int i = 0; i < nreg; ++i) if (i != iY && i != iStale) { o = i; break; }
    SpecArgs<T> sa = blank_args();
    sa.mode = STEP_CALCN; sa.Nout = reg[o];
    rank_barrier();
    t_eval = (double)t_;
    rhs(reg[iY], sa, true);
    sync_all();
    check_flags();
    absorb_red();
    iStale = iY;   // calcN! left vars.* = irfft(sol): the stale view is the state itself now
    const size_t fe = (P_ > 1) ? (size_t)nkr * Kyl * nz : (size_t)nkr * ny * nz;
    for (int f = 0; f < F; ++f) unpack_to_host(reg[o] + f * cf, (C*)p + (size_t)f * fe);
  }

  // ---- API boundary: real / spectral fields ------------------------------------------------
  void check_field(int f) const {
    if (f < 0 || f >= F) throw Err{MHDF_ERR_INVALID, "field index out of range"};
  }
  // host real field (this rank's z slab) -> compact spectral field at dst; leaves the sum / max of f^2 in red_h
  void real_to_compact(const void* p, C* dst, int emhd_slot) {
    T* re = reinterpret_cast<T*>(R);
    const size_t n = (size_t)nx * ny * nzl;
    CK(cudaMemcpyAsync(re, p, n * sizeof(T), cudaMemcpyDefault, st));
    if (emhd_slot >= 0)   // vars.b* <- the (undealiased) IC real field (IC.jl:86-90)
      CK(cudaMemcpyAsync(bst + (size_t)emhd_slot * n, re, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    XArgs<T> xa = xargs();
    xa.real_io = re; xa.out = Q; xa.in = nullptr;
    red_reset();
    xa.red = red_d;
    launch_xplain<-1>(xa);
    finish_red();
    from_xlayout(Q, R, dst, 1);
    sync_all();
  }
  void set_real(int field, const void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    real_to_compact(p, reg[iY] + field * cf, phys == MHDF_EMHD ? field : -1);
    // vars.* <- the copied-in field (copyto!(prob_ui, ui), IC.jl:74,88): the stale view follows sol, and so do its statistics
    if (iStale >= 0 && iStale != iY) {
      CK(cudaMemcpyAsync(reg[iStale] + field * cf, reg[iY] + field * cf, (size_t)cf * sizeof(C), cudaMemcpyDeviceToDevice, st));
      sync_all();
    }
    const int slot = (phys == MHDF_EMHD) ? 3 + field : field;
    st_sum[slot] = red_h->sumsq[0];
    st_max[slot] = red_max(0);
  }
  void set_forcing(int field, const void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    if (p == nullptr) { fmask &= ~(1u << field); return; }
    if (!force) force = dalloc<C>((size_t)F * cf);
    real_to_compact(p, force + field * cf, -1);
    fmask |= 1u << field;
  }
  void set_forcing_spectral(int field, const void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    if (p == nullptr) { fmask &= ~(1u << field); return; }
    if (!force) force = dalloc<C>((size_t)F * cf);
    const int nyh = (P_ > 1) ? Kyl : ny;
    const size_t n = (size_t)nkr * nyh * nz;
    CK(cudaMemcpyAsync(R, p, n * sizeof(C), cudaMemcpyDefault, st));
    MHDF_LAUNCH((k_pack<T>), pack_grid(), 256, 0, st, R, force + field * cf, nkr, nyh, nz, Kx, Kxp, by, bz, 0, P_ > 1);
    ++launches;
    CK(cudaGetLastError());
    sync_all();
    fmask |= 1u << field;
  }
  // ---- calcF! as an arbitrary host function (pgen.jl:231-234) ------------------------------------------------------------------
  // Called at the beginning of every RHS evaluation, when no transform buffer is live: the callback's own API calls (get_spectral,
  // set_forcing...) stage through R / Q.  Slab runs: the un-aliased-buffer argument of the pipelined exchange does not cover API
  // traffic in the middle of a step, so the callback is bracketed by cross-rank barriers with every rank's pushes drained.
  mhdf_forcing_fn fcb = nullptr;
  void* fcb_user = nullptr;
  const C* stage_in = nullptr;   // `sol` of the evaluation in flight (MHDF_STAGE), valid inside the callback only
  double t_eval = 0;             // its stage time
  bool in_cb = false;
  void set_forcing_callback(mhdf_forcing_fn fn, void* user) override {
    if (fn != nullptr && (a99.variant != A99_OFF || nd_on)) throw Err{MHDF_ERR_STATE, "one calcF! per problem: a forcing callback cannot be combined with the built-in driving"};
    if (fn != nullptr && cfg.stepper == MHDF_HM89) throw Err{MHDF_ERR_STATE, "HM89 with a forcing function is not supported (HM89.jl:182-196)"};
    fcb = fn; fcb_user = user;
  }
  void forcing_callback(const C* Sin) {
    if (fcb == nullptr || in_cb) return;
    sync_all();
    check_flags();
    if (P_ > 1) { rank_barrier(); sync_all(); }
    stage_in = Sin;
    in_cb = true;
    const int rc = fcb(fcb_user, t_eval);
    in_cb = false;
    stage_in = nullptr;
    if (P_ > 1) rank_barrier();
    if (rc != 0) throw Err{MHDF_ERR_STATE, "the forcing callback reported an error"};
  }
  void set_forcing_a99(const mhdf_a99* p) override {
    if (p == nullptr) { a99.variant = A99_OFF; return; }
    if (p->variant != A99_HOST && p->variant != A99_GPU) throw Err{MHDF_ERR_INVALID, "a99.variant must be MHDF_A99_HOST or MHDF_A99_GPU"};
    if (!(p->sigma2 > 0) || !std::isfinite(p->amp) || !std::isfinite(p->kf) || !(p->b != 0)) throw Err{MHDF_ERR_INVALID, "a99: need sigma2 > 0, b != 0, finite amp and kf"};
    a99.variant = p->variant; a99.nkr = nkr;
    a99.amp = (T)p->amp; a99.kf = (T)p->kf; a99.sig2 = (T)p->sigma2; a99.b = (T)p->b;
    a99.itanh = (T)(1.0 / std::tanh((double)(T)p->b * 3.14159265358979323846 / 2));
    a99.seed_lo = (unsigned)p->seed; a99.seed_hi = (unsigned)(p->seed >> 32);
    a99_call = p->call;
  }
  unsigned long long a99_calls() const override { return a99_call; }
  // params.χ / params.U₀x ... / params.B₀x ... (datastructure.jl:80-81,94-95): which = 0 chi, 1..3 U0, 4..6 B0 (MHD)
  void set_vp_field(int which, const void* p) override {
    if (!vp_on) throw Err{MHDF_ERR_STATE, "the problem was created without VP_method"};
    if (which < 0 || which > F) throw Err{MHDF_ERR_INVALID, "VP field index out of range (0 chi, 1..3 U0, 4..6 B0)"};
    CK(cudaSetDevice(cfg.device));
    const size_t n = (size_t)nx * ny * nzl;
    CK(cudaMemcpyAsync(vp_d + (size_t)which * n, p, n * sizeof(T), cudaMemcpyDefault, st));
    sync_all();
  }
  // SetUpND!(prob, P, fx, fy, fz) (pgen/NegativeDamping.jl:14-21): the power P and the three real profiles f_i
  void set_forcing_nd(double Pw, const void* fx, const void* fy, const void* fz) override {
    if (!nd_on) {
      if (cfg.nd != 0) return;   // HD / EMHD problem created with NDForceDriving!: the forcing never reaches N, like the reference
      throw Err{MHDF_ERR_STATE, "the problem was created without NDForceDriving! (mhdf_config.nd)"};
    }
    if (a99.variant != A99_OFF || fmask) throw Err{MHDF_ERR_STATE, "one calcF! per problem: NDForceDriving! cannot be combined with another forcing"};
    CK(cudaSetDevice(cfg.device));
    const size_t n = (size_t)nx * ny * nzl;
    const void* f[3] = {fx, fy, fz};
    for (int i = 0; i < 3; ++i)
      if (f[i]) CK(cudaMemcpyAsync(nd_d + (size_t)i * n, f[i], n * sizeof(T), cudaMemcpyDefault, st));
    sync_all();
    nd_P = Pw;
  }
  // DivVCorrection! (group 0) / DivBCorrection! (group 1), Solver/VPSolver.jl:61-137: project sol, then refresh the
  // real-space vars of that group (ldiv!(vars.bx, rfftplan, deepcopy(bxh)) ...) = the stale view and its statistics.
  void div_correction(int group) override {
    CK(cudaSetDevice(cfg.device));
    int f0;
    if (group == 0) { if (phys == MHDF_EMHD) throw Err{MHDF_ERR_INVALID, "DivVCorrection!: the EMHD state has no velocity"}; f0 = 0; }
    else if (group == 1) { if (phys == MHDF_HD) throw Err{MHDF_ERR_INVALID, "DivBCorrection!: the HD state has no magnetic field"}; f0 = (phys == MHDF_EMHD) ? 0 : 3; }
    else throw Err{MHDF_ERR_INVALID, "group must be 0 (velocity) or 1 (magnetic field)"};
    C* Y = reg[iY] + (size_t)f0 * cf;
    MHDF_LAUNCH((k_divclean<T>), spec_grid(), 256, 0, st, geom(), Y);
    ++launches;
    CK(cudaGetLastError());
    refresh_vars(f0);
  }
  // vars.* of the three fields starting at state field f0 <- c2r of sol (ldiv!(vars.bx, rfftplan, deepcopy(bxh)) ...,
  // VPSolver.jl:95-97,133-135): the stale register, EMHD's real-space b, and the CFL maxima / energies of those fields.
  void refresh_vars(int f0) {
    C* Y = reg[iY] + (size_t)f0 * cf;
    if (iStale >= 0 && iStale != iY)
      CK(cudaMemcpyAsync(reg[iStale] + (size_t)f0 * cf, Y, 3 * (size_t)cf * sizeof(C), cudaMemcpyDeviceToDevice, st));
    T* re = reinterpret_cast<T*>(R);
    const size_t n = (size_t)nx * ny * nzl;
    for (int i = 0; i < 3; ++i) {
      to_xlayout(Y + (size_t)i * cf, 1);
      XArgs<T> xa = xargs();
      xa.real_io = re; xa.in = Q; xa.out = nullptr;
      red_reset();
      xa.red = red_d;
      launch_xplain<+1>(xa);
      if (phys == MHDF_EMHD) CK(cudaMemcpyAsync(bst + (size_t)i * n, re, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
      finish_red();
      sync_all();
      const int slot = (phys == MHDF_EMHD) ? 3 + i : f0 + i;
      st_sum[slot] = red_h->sumsq[0];
      st_max[slot] = red_max(0);
    }
  }
  // ScaleDecomposition (mode 0) / VectorPotential (mode 1) of a vector field of the state (utils/MHDAnalysis.jl:24-82, 129-174) and
  // the autocorrelation functions CF of its components (mode 2, utils/TurbStatTool.jl:67),
  // on the device: spectral mask / curl into a dead register, three inverse transforms, three real fields to `out3`.
  void analysis(int mode, int group, int which, double k1, double k2, void* out3) override {
    CK(cudaSetDevice(cfg.device));
    int f0;
    if (group == 0) { if (phys == MHDF_EMHD) throw Err{MHDF_ERR_INVALID, "the EMHD state has no velocity"}; f0 = 0; }
    else if (group == 1) { if (phys == MHDF_HD) throw Err{MHDF_ERR_INVALID, "the HD state has no magnetic field"}; f0 = (phys == MHDF_EMHD) ? 0 : 3; }
    else throw Err{MHDF_ERR_INVALID, "group must be 0 (velocity) or 1 (magnetic field)"};
    if (mode < 0 || mode > 2) throw Err{MHDF_ERR_INVALID, "analysis mode must be 0 (scale decomposition), 1 (vector potential) or 2 (autocorrelation)"};
    int o = -1;
    const int nreg = nreg_();
    for (int i = 0; i < nreg; ++i) if (i != iY && i != iStale) { o = i; break; }
    rank_barrier();
    if (mode == 2) { gather_mirror(source(which)); wait_mirror(); }   // the power spectrum needs the symmetrised kr = 0 plane
    MHDF_LAUNCH((k_analysis<T>), spec_grid(), 256, 0, st, geom(), source(which), f0, reg[o], mode, (T)k1, (T)k2);
    ++launches;
    CK(cudaGetLastError());
    T* re = reinterpret_cast<T*>(R);
    const size_t n = (size_t)nx * ny * nzl;
    for (int i = 0; i < 3; ++i) {
      to_xlayout(reg[o] + (size_t)i * cf, 1);
      XArgs<T> xa = xargs();
      xa.real_io = re; xa.in = Q; xa.out = nullptr;
      launch_xplain<+1>(xa);
      CK(cudaMemcpyAsync(reinterpret_cast<T*>(out3) + (size_t)i * n, re, n * sizeof(T), cudaMemcpyDefault, st));
      sync_all();
    }
  }
  // DivFreeSpectraMap (utils/IC.jl:130-179) followed by SetUpProblemIC! (IC.jl:41-109) for one vector field, on the device:
  // the random-phase power-law spectra go straight onto the retained modes of sol (the reference's irfft -> copy -> rfft round
  // trip is the identity on them: the map is zero on the kr = 0 plane and dealiased), then vars.* are refreshed from sol.
  void set_random_phase(int group, unsigned long long seed, double k0, double Pw, double k_peak) override {
    CK(cudaSetDevice(cfg.device));
    int f0;
    if (group == 0) { if (phys == MHDF_EMHD) throw Err{MHDF_ERR_INVALID, "random-phase field: the EMHD state has no velocity (IC.jl:69)"}; f0 = 0; }
    else if (group == 1) { if (phys == MHDF_HD) throw Err{MHDF_ERR_INVALID, "random-phase field: the HD state has no magnetic field"}; f0 = (phys == MHDF_EMHD) ? 0 : 3; }
    else throw Err{MHDF_ERR_INVALID, "group must be 0 (velocity) or 1 (magnetic field)"};
    if (!(Pw > 0) || !std::isfinite(k0) || !std::isfinite(k_peak)) throw Err{MHDF_ERR_INVALID, "random-phase field: need P > 0 and finite k0, k_peak"};
    DfsmArgs<T> q;
    q.nkr = nkr; q.ny = ny; q.nz = nz;
    q.dkx = 2.0 * M_PI / cfg.Lx; q.dky = 2.0 * M_PI / cfg.Ly; q.dkz = 2.0 * M_PI / cfg.Lz;
    q.k0 = (T)k0; q.kpeak = (T)k_peak; q.amp = (T)0;
    q.seed_lo = (unsigned)seed; q.seed_hi = (unsigned)(seed >> 32);
    CK(cudaMemsetAsync(diag_d, 0, 8 * sizeof(double), st));
    MHDF_LAUNCH((k_dfsm_norm<T>), pack_grid(), 256, 0, st, q, diag_d);   // every rank sums the whole array: no collective needed
    ++launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(diag_h, diag_d, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync_all();
    const double intF = diag_h[0];
    if (!(intF > 0) || !std::isfinite(intF)) throw Err{MHDF_ERR_INVALID, "random-phase field: the spectrum k^k0 (k >= k_peak) is empty on this grid"};
    const double dx = cfg.Lx / nx, dy = cfg.Ly / ny, dz = cfg.Lz / nz;
    q.amp = (T)std::sqrt(Pw * 3 * (cfg.Lx / dx) * (cfg.Ly / dy) * (cfg.Lz / dz) / intF * (1 / dx / dy / dz));   // IC.jl:150
    MHDF_LAUNCH((k_dfsm_fill<T>), spec_grid(), 256, 0, st, geom(), q, reg[iY] + (size_t)f0 * cf);
    ++launches;
    CK(cudaGetLastError());
    refresh_vars(f0);
  }
  const C* source(int which) const {
    if (which == MHDF_STAGE && stage_in != nullptr) return stage_in;
    if (which == MHDF_STALE && iStale >= 0) return reg[iStale];
    return reg[iY];
  }
  void get_real(int field, int which, void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    to_xlayout(source(which) + field * cf, 1);
    T* re = reinterpret_cast<T*>(R);
    XArgs<T> xa = xargs();
    xa.real_io = re; xa.in = Q; xa.out = nullptr;
    launch_xplain<+1>(xa);
    CK(cudaMemcpyAsync(p, re, (size_t)nx * ny * nzl * sizeof(T), cudaMemcpyDefault, st));
    sync_all();
  }
  void set_spectral(int field, const void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    const int nyh = (P_ > 1) ? Kyl : ny;     // slab runs exchange the local compact ky rows directly
    const size_t n = (size_t)nkr * nyh * nz;
    CK(cudaMemcpyAsync(R, p, n * sizeof(C), cudaMemcpyDefault, st));
    MHDF_LAUNCH((k_pack<T>), pack_grid(), 256, 0, st, R, reg[iY] + field * cf, nkr, nyh, nz, Kx, Kxp, by, bz, 0, P_ > 1);
    ++launches;
    CK(cudaGetLastError());
    sync_all();
  }
  int pack_grid() const {
    long long b = ((long long)nkr * ny * nz + 255) / 256;
    long long cap = (long long)nsm * 16;
    return (int)(b < cap ? b : cap);
  }
  void unpack_to_host(const C* comp, C* host) {
    const int nyh = (P_ > 1) ? Kyl : ny;
    const size_t n = (size_t)nkr * nyh * nz;
    MHDF_LAUNCH((k_pack<T>), pack_grid(), 256, 0, st, R, const_cast<C*>(comp), nkr, nyh, nz, Kx, Kxp, by, bz, 1, P_ > 1);
    ++launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, R, n * sizeof(C), cudaMemcpyDefault, st));
    sync_all();
  }
  void get_spectral(int field, int which, void* p) override {
    check_field(field);
    CK(cudaSetDevice(cfg.device));
    unpack_to_host(source(which) + field * cf, (C*)p);
  }

  // ---- clock, CFL, diagnostics ----------------------------------------------------------------
  void set_dt(double dt) override { dt_ = (T)dt; }
  void set_clock(double t, long long s) override { t_ = (T)t; step_ = s; }
  void get_clock(double* t, double* dt, long long* s) const override {
    if (t) *t = (double)t_;
    if (dt) *dt = (double)dt_;
    if (s) *s = step_;
  }
  void cfl_dt(double coef, double t_diff, double* dt) override {
    // integrator.jl:158-198 on the stale maxima (vars.* of the last RHS evaluation)
    double vmax = std::sqrt(std::fmax(st_max[0], std::fmax(st_max[1], st_max[2])));   // u, or curl B for EMHD
    if (phys != MHDF_HD) {
      const double va = std::sqrt(std::fmax(st_max[3], std::fmax(st_max[4], st_max[5])));
      vmax = std::fmax(vmax, va);
    }
    const double dx = cfg.Lx / nx, dy = cfg.Ly / ny, dz = cfg.Lz / nz;
    double dl = std::fmin(dx, std::fmin(dy, dz));
    if (phys == MHDF_EMHD) dl = dl * dl;
    double d = coef * dl / vmax;
    if (!(d < t_diff)) d = t_diff;
    dt_ = (T)d;
    if (dt) *dt = (double)dt_;
  }
  double dV() const {
    // ProbDiagnostic: dV = diff(x)[1]*diff(y)[1]*diff(z)[1] with x = range(T(x0), step=T(dx)) (UserInterface.jl:66-67)
    return (double)(T)(cfg.Lx / nx) * (double)(T)(cfg.Ly / ny) * (double)(T)(cfg.Lz / nz);
  }
  void run_diag(int which) {
    const C* src = source(which);
    gather_mirror(src);
    CK(cudaMemsetAsync(diag_d, 0, 8 * sizeof(double), st));
    const int has_u = (phys != MHDF_EMHD), has_b = (phys != MHDF_HD), boff = (phys == MHDF_EMHD) ? 0 : 3;
    wait_mirror();
    MHDF_LAUNCH((k_diag<T>), spec_grid(), 256, 0, st, geom(), src, has_u, has_b, boff, 1.0 / ((double)nx * ny * nz), diag_d);
    ++launches;
    CK(cudaGetLastError());
    if (P_ > 1) { order(st, sc); NK(g_nccl.AllReduce(diag_d, diag_d, 8, ncclFloat64, ncclSum, comm, sc)); order(sc, st); }
    CK(cudaMemcpyAsync(diag_h, diag_d, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync_all();
  }
  void energy(int which, double* KE, double* ME) override {
    CK(cudaSetDevice(cfg.device));
    double ke, me;
    if (which == MHDF_STALE) {
      if (phys == MHDF_EMHD) { ke = 0; me = (st_sum[3] + st_sum[4] + st_sum[5]) * dV(); }
      else { ke = (st_sum[0] + st_sum[1] + st_sum[2]) * dV(); me = (st_sum[3] + st_sum[4] + st_sum[5]) * dV(); }
      if (phys == MHDF_HD) me = 0;
    } else {
      run_diag(MHDF_FRESH);
      ke = diag_h[0] * dV(); me = diag_h[1] * dV();
    }
    if (std::isnan(ke) || std::isnan(me)) throw Err{MHDF_ERR_NONFINITE, "detected NaN! Quit the simulation right now."};
    if (KE) *KE = ke;
    if (ME) *ME = me;
  }
  void helicity(double* Hk, double* Hm, double* Hc) override {
    CK(cudaSetDevice(cfg.device));
    run_diag(MHDF_FRESH);
    const double dv = (cfg.Lx / nx) * (cfg.Ly / ny) * (cfg.Lz / nz);
    if (Hk) *Hk = diag_h[2] * dv;
    if (Hm) *Hm = diag_h[3];
    if (Hc) *Hc = diag_h[4] * dv;
  }
  void spectrum(int field, double* Pk, int nbins) override {
    check_field(field);
    if (nbins <= 0 || nbins > 4096) throw Err{MHDF_ERR_INVALID, "nbins must be in 1..4096"};
    CK(cudaSetDevice(cfg.device));
    if (spec_cap < nbins) {
      if (spec_d) cudaFree(spec_d);
      spec_d = nullptr;
      CK(cudaMalloc(&spec_d, nbins * sizeof(double)));
      spec_cap = nbins;
    }
    CK(cudaMemsetAsync(spec_d, 0, nbins * sizeof(double), st));
    gather_mirror(reg[iY]);
    wait_mirror();
    MHDF_LAUNCH((k_spectrum<T>), spec_grid(), 256, nbins * sizeof(double), st, geom(), reg[iY], field, spec_d, nbins);
    ++launches;
    CK(cudaGetLastError());
    if (P_ > 1) { order(st, sc); NK(g_nccl.AllReduce(spec_d, spec_d, nbins, ncclFloat64, ncclSum, comm, sc)); order(sc, st); }
    CK(cudaMemcpyAsync(Pk, spec_d, nbins * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync_all();
  }
  void stale_stats(double* mx, double* sm) const override {
    for (int i = 0; i < 6; ++i) { if (mx) mx[i] = st_max[i]; if (sm) sm[i] = st_sum[i]; }
  }
  void profile(int enable) override {
    if (prof && !enable) prof_collect();
    prof = enable != 0;
    if (enable) for (int i = 0; i < KC_COUNT; ++i) { prof_ms[i] = 0; prof_cnt[i] = 0; }
  }
  void profile_get(double* ms, long long* cnt, int n) override {
    prof_collect();
    for (int i = 0; i < n && i < KC_COUNT; ++i) { if (ms) ms[i] = prof_ms[i]; if (cnt) cnt[i] = prof_cnt[i]; }
  }
  long long launch_count() const override { return launches; }
  void info(int* nf, int* kx, int* kxp, int* ky, int* kz, long long* bytes) const override {
    if (nf) *nf = F;
    if (kx) *kx = Kx;
    if (kxp) *kxp = Kxp;
    if (ky) *ky = Ky;
    if (kz) *kz = Kz;
    if (bytes) *bytes = bytes_dev;
  }
};


// factories defined in solver_f32.cu / solver_f64.cu
mhdf_handle* mhdf_make_solver_f32(const mhdf_config& c);
mhdf_handle* mhdf_make_solver_f64(const mhdf_config& c);
