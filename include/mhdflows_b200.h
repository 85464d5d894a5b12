/* mhdflows_b200.h -- C ABI of the B200-native MHDFlows hot path.
 *
 * Drop-in boundary for MHDFlows.jl's 3D periodic pseudospectral right-hand side and
 * RK4 / LSRK54 time step (incompressible HD, MHD, electron MHD).  A Julia host binds these
 * with plain `ccall` (see INTEGRATION.md and julia/MHDFlowsB200.jl); the Python ctypes mirror
 * is mhdflows_jl_b200/.  Plain pointers and sizes only; every function returns 0 on success or
 * a negative mhdf_status; nothing throws across the boundary.
 *
 * Arrays use the reference's column-major layouts so Julia passes `pointer(A)` unchanged:
 *   real field      (nx, ny, nz)              x fastest
 *   spectral field  (nx/2+1, ny, nz)          kx fastest, interleaved (re, im)
 * Field ids follow params.*_ind - 1 (src/Structure/datastructure.jl:88-90):
 *   HD: ux,uy,uz = 0,1,2   MHD: ux,uy,uz,bx,by,bz = 0..5   EMHD: bx,by,bz = 0,1,2
 *
 * The handle owns all device memory, streams and communicators; caller pointers are borrowed
 * for the duration of a call.  One host thread per handle.
 * Field pointers (`host_real`, `host_spec`) may be host memory OR device memory of the handle's GPU
 * (unified addressing, cudaMemcpyDefault): a Julia host passes `pointer(A)` of an Array or of a CuArray
 * alike -- the reference's `vars.*` / `sol` live on the device (datastructure.jl:58-108).
 *
 * file:line citations are into the reference tree (MHDFlows.jl).
 */
#ifndef MHDFLOWS_B200_H
#define MHDFLOWS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mhdf_handle mhdf_handle;

typedef enum {
  MHDF_OK = 0,
  MHDF_ERR_INVALID = -1,   /* bad argument / unsupported configuration (Julia: error(...), pgen.jl:98-105) */
  MHDF_ERR_CUDA = -2,      /* CUDA runtime failure, including "no CUDA device" */
  MHDF_ERR_NCCL = -3,
  MHDF_ERR_NONFINITE = -4, /* NaN detected (UserInterface.jl:71,79 "detected NaN! Quit the simulation right now.") */
  MHDF_ERR_STATE = -5      /* call not valid in the current state */
} mhdf_status;

enum { MHDF_F32 = 0, MHDF_F64 = 1 };
enum { MHDF_HD = 0, MHDF_MHD = 1, MHDF_EMHD = 2 };          /* B_field / EMHD flags, pgen.jl:129-150 */
enum { MHDF_RK4 = 0, MHDF_LSRK54 = 1, MHDF_HM89 = 2 };       /* stepper = "RK4" | "LSRK54" | "HM89" (EMHD only), Problems.jl:123-128 */
enum { MHDF_FRESH = 0, MHDF_STALE = 1, MHDF_STAGE = 2 };     /* which view: true state, the reference's `vars.*` = c2r of the
                                                               last stage input, or -- inside a forcing callback only --
                                                               the `sol` argument of the evaluation in flight */

/* Keyword arguments of Problem(dev; ...) (pgen.jl:64-95) that reach the hot path. */
typedef struct {
  int nx, ny, nz;          /* powers of two, 16..1024 */
  double Lx, Ly, Lz;
  double nu, eta;          /* params.nu, params.eta */
  int n_nu;                /* params.n_nu (hyperviscosity ADDS on top of viscosity when > 1, MHDSolver.jl:94-99) */
  double dt;               /* clock.dt */
  int physics;             /* MHDF_HD | MHDF_MHD | MHDF_EMHD */
  int stepper;             /* MHDF_RK4 | MHDF_LSRK54 | MHDF_HM89 (HM89TimeStepper, timestepper/HM89.jl; needs physics = MHDF_EMHD) */
  int dtype;               /* MHDF_F32 | MHDF_F64  (T = Float32 default, pgen.jl:90) */
  int device;              /* CUDA device ordinal */
  /* slab decomposition over `nranks` processes (one GPU each); rank 0..nranks-1.  nranks = 1: single GPU. */
  int rank, nranks;
  const void* nccl_id;     /* 128-byte ncclUniqueId from mhdf_nccl_unique_id on rank 0, shared by the host; NULL if nranks == 1 */
  int vp;                  /* VP_method (pgen.jl:84): volume-penalisation terms in the HD / MHD right-hand side (VPSolver.jl:21-59) */
  int nd;                  /* calcF = NDForceDriving! (pgen/NegativeDamping.jl): negative-damping forcing, set with mhdf_set_forcing_nd */
} mhdf_config;

/* Problem(...) constructor / finaliser (pgen.jl:64-127, Problems.jl:118-140). */
int mhdf_create(const mhdf_config* cfg, mhdf_handle** out);
int mhdf_destroy(mhdf_handle* h);
const char* mhdf_last_error(const mhdf_handle* h);   /* h may be NULL: error of the last failed mhdf_create */
int mhdf_nccl_unique_id(void* id128);

/* Slab runs only: peer-memory exchange.  Each rank exports a blob (CUDA IPC handles of its two exchange buffers), the
 * host gathers the nranks blobs (rank order) and hands them to every rank.  Afterwards the global transposes are
 * copy-engine pushes into the peers' HBM over NVLink instead of NCCL send/recv.  No reference counterpart
 * (the reference is single-device, README.md:40-41). */
int mhdf_ipc_blob_size(const mhdf_handle* h);
int mhdf_ipc_export(mhdf_handle* h, void* blob);
int mhdf_ipc_import(mhdf_handle* h, const void* all_blobs);

/* SetUpProblemIC! (utils/IC.jl:41-109): copy a real field in and r2c it into sol[:, :, :, field]. */
int mhdf_set_real(mhdf_handle* h, int field, const void* host_real);
/* vars.ux ... (c2r on demand).  which = MHDF_FRESH | MHDF_STALE. */
int mhdf_get_real(mhdf_handle* h, int field, int which, void* host_real);
/* prob.sol[:, :, :, field] in the (nx/2+1, ny, nz) layout.  Dealiased modes read back as zero: the state is
 * stored on the modes kept by dealias!() only (equals the reference's sol after its next dealias!, pgen.jl:155). */
int mhdf_set_spectral(mhdf_handle* h, int field, const void* host_spec);
int mhdf_get_spectral(mhdf_handle* h, int field, int which, void* host_spec);

/* Constant forcing: the reference's `calcF!` hook (pgen.jl:231-234) for a time-independent forcing given in real space,
 * e.g. N97ForceDriving! (pgen/TaylorGreenDynamo.jl:12-34): N[:, :, :, field] += rfft(F) on every RHS evaluation.
 * host_real = NULL removes the forcing of that field.  Like the reference, it acts on the MHD path only: HDcalcN! adds
 * the forcing BEFORE the advection zeroes N (pgen.jl:176-178, HDSolver.jl:55) and EMHDcalcN! never calls it. */
int mhdf_set_forcing(mhdf_handle* h, int field, const void* host_real);
/* Arbitrary calcF! closures (pgen.jl:231-234: `params.calcF!(N, sol, t, clock, vars, params, grid)`, any host function): the
 * library calls `fn(user, t)` on the host thread of mhdf_step / mhdf_calcN at the beginning of EVERY right-hand-side evaluation
 * (t = the stage time FourierFlows passes).  Inside, the host may read the evaluation's `sol` with mhdf_get_spectral / mhdf_get_real
 * (which = MHDF_STAGE) and defines what the evaluation adds to N with mhdf_set_forcing_spectral (or mhdf_set_forcing); a non-zero
 * return aborts the step with MHDF_ERR_STATE.  Additive forcings only (N += F: every forcing the reference ships); `vars.*` seen
 * by the callback are the stale ones.  Costs a device synchronisation (and, on slabs, two cross-rank barriers) per evaluation: the
 * compatibility path for user code, not the fast path -- the reference's own forcings are built in.  fn = NULL removes it. */
typedef int (*mhdf_forcing_fn)(void* user, double t);
int mhdf_set_forcing_callback(mhdf_handle* h, mhdf_forcing_fn fn, void* user);
/* forcing of one field as a spectral array in the layout of mhdf_set_spectral (dealiased modes are ignored); NULL removes it */
int mhdf_set_forcing_spectral(mhdf_handle* h, int field, const void* fhat);

/* Random solenoidal driving (Alvelius 1999): the reference's `calcF!` = A99ForceDriving!.  Every RHS evaluation adds
 *   N_u += amp * Fk(k) * (e^{i th1} g_i e1(k) + e^{i th2} g_j e2(k)),   Fk = sqrt(exp(-(k-kf)^2/sigma2) / 2pi) / k
 * with fresh uniform random th1, th2, Phi per mode.  variant selects which of the reference's two implementations is
 * restated (basis vectors, real or complex Phi, treatment of the kr = 0 plane):
 *   MHDF_A99_HOST  top-level A99ForceDriving! with the tables of SetUpFk (pgen/A99ForceDriving.jl:33-60, 93-127)
 *   MHDF_A99_GPU   module A99GPU (pgen/A99ForceDriving_GPU.jl:49-130)
 * amp = usr_vars.A times the normalisation A of SetUpFk (the host computes it, as the reference does).
 * Random numbers: Philox4x32-10, key = seed, counter = (index of the mode in the (nx/2+1, ny, nz) array, number of the RHS
 * evaluation): reproducible, independent of the number of GPUs; `call` is the evaluation number to start from (restart).
 * Acts on the MHD path only, like mhdf_set_forcing.  p = NULL removes the driving. */
enum { MHDF_A99_HOST = 1, MHDF_A99_GPU = 2 };
typedef struct {
  int variant;
  double amp, kf, sigma2, b;
  unsigned long long seed, call;
} mhdf_a99;
int mhdf_set_forcing_a99(mhdf_handle* h, const mhdf_a99* p);
int mhdf_forcing_a99_calls(const mhdf_handle* h, unsigned long long* calls);

/* Negative-damping forcing: the reference's `calcF!` = NDForceDriving! with SetUpND!(prob, P, fx, fy, fz)
 * (pgen/NegativeDamping.jl:14-45).  Every RHS evaluation adds  N_ui += A rfft(f_i u_i),  A = P / (sum_i sum |u_i^2 f_i| dV),  with the
 * u of that evaluation: the products ride along the forward transforms of the fused x kernel, the sum is a reduction of the same
 * kernel (all-reduced over the ranks of a slab run).  Needs mhdf_config.nd at creation; acts on the MHD path only, like every
 * forcing of the reference; a NULL profile pointer keeps the previous profile; P = 0 switches the forcing off. */
int mhdf_set_forcing_nd(mhdf_handle* h, double P, const void* fx, const void* fy, const void* fz);

/* Volume penalisation (Problem(...; VP_method = true)): the real fields params.χ (which = 0), params.U₀x,U₀y,U₀z (1..3) and, for
 * MHD, params.B₀x,B₀y,B₀z (4..6) (datastructure.jl:80-81,94-95; SetUpProblemIC! keywords U₀x ..., IC.jl:93-106).  Every RHS
 * evaluation then adds  N_a += -sum_j (delta_aj - k_j k_a / k^2) F[chi/eta (u_j - U0_j)],  eta = clock.dt * 13/7, to the
 * velocity equation and the same with (b, B0) to the induction equation (VPSolver.jl:21-59).  All fields start as zero. */
int mhdf_set_vp_field(mhdf_handle* h, int which, const void* host_real);

/* DivVCorrection! (group 0) / DivBCorrection! (group 1) (Solver/VPSolver.jl:61-137): sol_i -= k_i (k . sol) / k^2 on the
 * three fields of the group, then the real-space `vars` of that group are refreshed from the corrected sol (stale view,
 * CFL maxima, energies). */
int mhdf_div_correction(mhdf_handle* h, int group);

/* DivFreeSpectraMap(grid; k_peak, P, k0) (utils/IC.jl:130-179) + SetUpProblemIC! (IC.jl:41-109) for the velocity (group 0) or
 * the magnetic field (group 1), entirely on the device (the reference builds the map on grid.device too):
 *   F^_i = A k^k0 e^{2 pi i theta(k)} e2_i(k),  e2 = (kx kz, ky kz, -kp^2) / (kp k),  zero on the kr = 0 plane and for k < k_peak,
 *   A = sqrt(3 P (Lx/dx)(Ly/dy)(Lz/dz) / sum(k^k0 / (k+1)^2) / dV),  then dealias!, and sol / vars.* of the group are set from it.
 * theta: Julia's rand stream cannot be reproduced; it is word 0 of the Philox4x32-10 block with key = seed and counter = index
 * of the mode in the (nx/2+1, ny, nz) array (independent of the number of GPUs; regenerated bit for bit by
 * oracle/forcing_oracle.py::PhiloxField for the parity tests). */
int mhdf_set_random_phase(mhdf_handle* h, int group, unsigned long long seed, double k0, double P, double k_peak);

/* stepforward! (timestepper/timestepper.jl:4-6): nsteps steps of clock.dt. */
int mhdf_step(mhdf_handle* h, int nsteps);
/* HM89TimeStepper only (timestepper/HM89.jl:61-84): how many fixed-point iterations the last step took and its last error norm
 * max |B^n - B^1| (the reference keeps both in locals of HM89substeps!); 0 and 0 for the explicit steppers. */
int mhdf_stepper_stats(mhdf_handle* h, long long* fixed_point_iters, double* last_error);
/* eqn.calcN!(N, sol, t, clock, vars, params, grid) (pgen.jl:153-181) on the current sol:
 * writes N as nfields spectral arrays (dealiased modes zero).  Refreshes the stale `vars` like the reference. */
int mhdf_calcN(mhdf_handle* h, void* host_N);
int mhdf_set_dt(mhdf_handle* h, double dt);
int mhdf_set_clock(mhdf_handle* h, double t, long long step);
int mhdf_get_clock(const mhdf_handle* h, double* t, double* dt, long long* step);

/* getCFL! (integrator.jl:158-198): dt = min(coef * dl / vmax, t_diff) from the stale real-space maxima; sets clock.dt. */
int mhdf_cfl_dt(mhdf_handle* h, double coef, double t_diff, double* dt_out);
/* ProbDiagnostic (utils/UserInterface.jl:65-86) without the sigdigits rounding: sum(u^2) dV, sum(b^2) dV. */
int mhdf_energy(mhdf_handle* h, int which, double* KE, double* ME);
/* sum (curl u).u dV (MHDAnalysis.jl:94-101), sum a.b with Coulomb-gauge a (MHDAnalysis.jl:113-117, no dV), sum u.b dV. */
int mhdf_helicity(mhdf_handle* h, double* Hk, double* Hm, double* Hc);
/* spectralline (MHDAnalysis.jl:237-255) of one state field: Pk[round(|k|)] += |f^|^2 over the half spectrum. */
int mhdf_spectrum(mhdf_handle* h, int field, double* Pk, int nbins);
/* stale per-field maxima of f^2 and sums of f^2 (6 each; EMHD: curl B then b) as used by getCFL!/ProbDiagnostic. */
int mhdf_stale_stats(const mhdf_handle* h, double* maxsq6, double* sumsq6);

/* On-device analysis of the state (utils/MHDAnalysis.jl), three real fields (x, y, z components, consecutive) to `out3`
 * (host or device memory):
 *   ScaleDecomposition(B1, B2, B3, grid; kf = [k1, k2]) (MHDAnalysis.jl:54-82): the part of the velocity (group 0) or of the
 *     magnetic field (group 1) with k1 <= |k| <= k2;
 *   VectorPotential(B1, B2, B3) (MHDAnalysis.jl:129-174): a = curl^-1 b in the Coulomb gauge, a^ = i (k x b^) / k^2;
 *   CF(V) = real(ifft(abs.(fft(V)).^2)) (utils/TurbStatTool.jl:67, WITHOUT its fftshift: zero lag at element 0) of each component of
 *     the velocity (group 0) or the magnetic field (group 1): the periodic autocorrelation from which the two-point structure
 *     functions SFC(V) = 2 (mean(V) - CF(V)) and the radial SF_2 1D (TurbStatTool.jl:72, 90-120) follow on the host.
 * which = MHDF_FRESH (sol) or MHDF_STALE (the reference's vars.*, what a user script would pass).  The reference applies these
 * to arbitrary arrays; here the input is the (dealiased) state, i.e. equal whenever the argument is band-limited. */
int mhdf_scale_decomposition(mhdf_handle* h, int group, int which, double k1, double k2, void* out3);
int mhdf_vector_potential(mhdf_handle* h, int which, void* out3);
int mhdf_correlation(mhdf_handle* h, int group, int which, void* out3);

/* Measurement helpers. */
int mhdf_step_timed(mhdf_handle* h, int nsteps, double* ms_total);   /* CUDA events on the library stream */
int mhdf_profile(mhdf_handle* h, int enable);                        /* per-kernel-class CUDA event timing */
/* classes: 0 z-inverse, 1 y-inverse, 2 fused x pass, 3 y-forward, 4 z-forward, 5 spectral update, 6 emhd derive, 7 exchange */
int mhdf_profile_get(mhdf_handle* h, double* ms_per_class, long long* launches_per_class, int nclasses);
long long mhdf_launch_count(const mhdf_handle* h);                   /* kernels launched by this handle so far */
int mhdf_info(const mhdf_handle* h, int* nfields, int* kx, int* kxp, int* ky, int* kz, long long* bytes_device);

#ifdef __cplusplus
}
#endif
#endif
