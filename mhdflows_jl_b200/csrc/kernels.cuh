// kernels.cuh -- sm_100a kernels of the pseudospectral RHS / time step.
//
// Data layout (all x/kx-fastest, matching the reference's column-major arrays):
//   compact spectral field  [Kz][Ky][Kxp]   only the modes kept by FourierFlows' dealias!()
//                                           (kx < Kx; ky,kz in two bands: [0,lo) and [n-hi,n))
//   after inverse z pass    [nz][Ky][Kxp]
//   after inverse y pass    [nz][ny][Kxp]   (x-pass input / output; real space is never stored)
//   real field (API only)   [nz][ny][nx]
// Kxp = Kx rounded up to 8 elements so every row starts 64-byte aligned; pad columns stay zero.
#pragma once
#include <type_traits>

#include "fft_core.cuh"

// MHDF_KEEP_PTR: make the compiler treat a pointer as an opaque 64-bit value from here on (so accesses become
// base + 32-bit offset, one IMAD.WIDE each).  MHDF_DYN_SMEM: the block's dynamic shared memory.
#ifdef MHDF_CPU_EMU
#define MHDF_KEEP_PTR(p) ((void)0)
#define MHDF_DYN_SMEM(type, name)
#else
#define MHDF_KEEP_PTR(p) asm volatile("" : "+l"(p))
#define MHDF_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace mhdf {

// Retained-band descriptor of one axis: full index n -> compact row, or -1 if dealiased.
struct Band {
  int n;    // full length
  int lo;   // rows [0, lo) kept        (non-negative wavenumbers 0..lo-1)
  int hi0;  // rows [hi0, n) kept       (negative wavenumbers -(n-hi0)..-1)
  __host__ __device__ int count() const { return lo + (n - hi0); }
  __host__ __device__ int row(int i) const { return i < lo ? i : (i >= hi0 ? i - (hi0 - lo) : -1); }
  // signed integer wavenumber of compact row c
  __host__ __device__ int wave(int c) const { return c < lo ? c : c - count(); }
  // compact row of signed wavenumber w, or -1
  __host__ __device__ int row_of_wave(int w) const {
    if (w >= 0) return w < lo ? w : -1;
    return (-w <= n - hi0) ? count() + w : -1;
  }
};

// ------------------------------------------------------------------------------------------------
// Strided axis pass (y or z): a block transforms TX adjacent columns of length N.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct PassArgs {
  const Cx<T>* in;
  Cx<T>* out;
  const Cx<T>* tw;          // twiddle table of length N: (cos, -sin)(2 pi i / N)
  int in_row, out_row;             // element stride between consecutive rows n of the FFT axis (< 2^31 within a field)
  long long in_outer, out_outer;   // stride of the outer (non-transformed, non-contiguous) axis
  long long in_field, out_field;   // stride between fields of the batch
  int inner;                       // contiguous elements per row to process
  // retained band of the pruned side: rows [0, lo) and [hi0, N) exist, stored contiguously (shift = hi0 - lo)
  int lo, hi0, shift;
  // BLK variants (slab-decomposed runs): one side lives in the blocked exchange layout [peer][field][z'][ky'][kx] that
  // the all-to-all moves as contiguous pieces.  Row r of that side sits at  q * blk_stride + (r - q * blk_rows) * row,
  // q = r / blk_rows computed as mulhi(r, blk_magic)  (blk_magic = ceil(2^32 / blk_rows), exact for r < 2^16).
  int blk_rows, blk_stride;
  unsigned blk_magic;
  // two-level variant (z-chunk pipelined slab runs, BLK 3 / 4): inside a peer's blk_rows rows, chunks of blk2_rows rows
  // are blk2_stride apart and the peer pieces blk_stride apart:  [chunk][peer][field][row in chunk][ky'][kx]
  int blk2_rows, blk2_stride;
  unsigned blk2_magic;
};
__device__ __forceinline__ int blk_off(int r, int rows, int stride, unsigned magic, int rowstride) {
  const int q = (int)__umulhi((unsigned)r, magic);
  return q * stride + (r - q * rows) * rowstride;
}
template <typename T>
__device__ __forceinline__ int blk_off2(int r, const PassArgs<T>& a, int rowstride) {
  const int q = (int)__umulhi((unsigned)r, a.blk_magic);
  const int rem = r - q * a.blk_rows;
  const int c = (int)__umulhi((unsigned)rem, a.blk2_magic);
  return c * a.blk2_stride + q * a.blk_stride + (rem - c * a.blk2_rows) * rowstride;
}

// predicated 8/16-byte global accesses (no branches, no speculative address use)
#ifdef MHDF_CPU_EMU
template <typename C> inline C ldg_pred(const C* p, bool ok) { C z; z.x = 0; z.y = 0; return ok ? *p : z; }
template <typename C> inline void stg_pred(C* p, C v, bool ok) { if (ok) *p = v; }
#else
// (L1::no_allocate on these loads was measured and rejected: the x pass re-reads X[M-k] through L1 -- +13 % on the x pass, no
// change on the strided passes; profiles/r02_c10_time1024.log)
__device__ __forceinline__ float2 ldg_pred(const float2* p, bool ok) {
  float2 v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
               "@p ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
               : "=f"(v.x), "=f"(v.y) : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ double2 ldg_pred(const double2* p, bool ok) {
  double2 v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
               "@p ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
               : "=d"(v.x), "=d"(v.y) : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ void stg_pred(float2* p, float2 v, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.f32 [%0], {%1, %2};\n\t}"
               :: "l"(p), "f"(v.x), "f"(v.y), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void stg_pred(double2* p, double2 v, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.f64 [%0], {%1, %2};\n\t}"
               :: "l"(p), "d"(v.x), "d"(v.y), "r"((int)ok) : "memory");
}
#endif

template <int N, int TX, int R1, typename C> struct PassIdx {
  // float2 with TX = 8: two rows share one 128-byte bank line; shift by 8 slots every R1 rows so
  // rows n and n+R1 (written together in the first exchange) land in different halves.
  static constexpr bool PAD = (sizeof(C) == 8 && TX == 8);
  static constexpr int SIZE = N * TX + (PAD ? (N / R1) * 8 : 0);   // exchange buffer; the twiddle table follows it
  int c;
  __device__ __forceinline__ int operator()(int n) const { return n * TX + c + (PAD ? (n / R1) * 8 : 0); }
};

// PIN: the input side is the pruned (band) side (inverse passes); otherwise the output side is (forward passes).
// BLK: 0 = plain strides, 1 = input side blocked, 2 = output side blocked, 3 / 4 = input / output side two-level blocked
// Float32: at most 64 registers per thread (launch bound) -- measured 13 % faster per 512^3 step than the 76-80
// registers the compiler takes otherwise, because one more block fits per SM.
// (32 points per thread -- the MHDF_PASS_E32 variant -- hold 64 data registers: budget 128)
constexpr int pass_minb(int threads, int tsize, int E = 16) {
  return tsize == 4 ? ((65536 / (threads * (E >= 32 ? 128 : 64))) > 16 ? 16 : (65536 / (threads * (E >= 32 ? 128 : 64)))) : 1;
}
// Arithmetic type of the strided passes: Float32 butterflies on the packed instructions (float2p, fft_core.cuh) -- measured on B200
// against the scalar forms, per pass: everything below 1024 points gains (256^3 step -2.5 %, 512^3 -0.9 %, profiles/r02_c11_*); at
// 1024 points the inverse passes gain 5-10 %, the forward y passes (8 columns per block) gain 9 % (37.3 -> 34.0 ms per 1024^3 step)
// and the forward z passes (16 columns) lose 2 % (profiles/r02_c17_time1024.log), so only those stay scalar.  MHDF_PASS_SCALAR:
// scalar everywhere, MHDF_PASS_FWD_PACKED: packed everywhere (A/B partners).  Data in memory is plain float2 / double2 either way.
template <typename T, int N, int DIR, int TX> struct PassCx { using type = Cx<T>; };
#ifndef MHDF_PASS_SCALAR
#ifdef MHDF_PASS_FWD_PACKED
template <int N, int DIR, int TX> struct PassCx<float, N, DIR, TX> { using type = float2p; };
#else
template <int N, int DIR, int TX> struct PassCx<float, N, DIR, TX> { using type = typename std::conditional<(N >= 1024 && DIR < 0 && TX >= 16), float2, float2p>::type; };
#endif
#endif
__device__ __forceinline__ float2p pass_in(float2 v, float2p) { return to_p(v); }
__device__ __forceinline__ float2 pass_out(float2p v) { return from_p(v); }
template <typename C> __device__ __forceinline__ C pass_in(C v, C) { return v; }
template <typename C> __device__ __forceinline__ C pass_out(C v) { return v; }

template <typename T, int N, int E, int TX, int DIR, bool PIN, int BLK, int MINB = pass_minb((N / E) * TX, (int)sizeof(T), E)>
__global__ void __launch_bounds__((N / E) * TX, MINB) k_pass(PassArgs<T> a) {
  using C = Cx<T>;                         // element type in memory
  using CA = typename PassCx<T, N, DIR, TX>::type; // element type of the arithmetic
  constexpr int Tn = N / E;
  constexpr int R1 = imin(E, N);
  MHDF_DYN_SMEM(unsigned char, smem_raw);
  CA* sm = reinterpret_cast<CA*>(smem_raw);
  const int c = threadIdx.x % TX;
  const int t = threadIdx.x / TX;
  const int col = blockIdx.x * TX + c;
  const bool valid = col < a.inner;
  const C* ip = a.in + ((long long)blockIdx.z * a.in_field + (long long)blockIdx.y * a.in_outer + col);
  C* op = a.out + ((long long)blockIdx.z * a.out_field + (long long)blockIdx.y * a.out_outer + col);
  const C* twp = a.tw;
  MHDF_KEEP_PTR(twp);

  // 32-bit unsigned element offsets from one materialised 64-bit base per side, so every access costs one IMAD.WIDE
  // (the empty asm keeps the compiler from re-deriving base + index * 8 from the kernel parameters per access)
  const unsigned irow = (unsigned)a.in_row, orow = (unsigned)a.out_row;
  MHDF_KEEP_PTR(ip);
  MHDF_KEEP_PTR(op);
  CA v[E];
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int n = t + Tn * m;
    if (PIN) {
      const bool hi = n >= a.hi0;
      const bool ok = valid && (n < a.lo || hi);
      const unsigned r = (unsigned)(n - (hi ? a.shift : 0));
      const unsigned off = (BLK == 1) ? (unsigned)blk_off((int)r, a.blk_rows, a.blk_stride, a.blk_magic, a.in_row)
                         : (BLK == 3) ? (unsigned)blk_off2<T>((int)r, a, a.in_row) : r * irow;
      v[m] = pass_in(ldg_pred(ip + off, ok), CA());
    } else {
      const unsigned off = (BLK == 1) ? (unsigned)blk_off(n, a.blk_rows, a.blk_stride, a.blk_magic, a.in_row)
                         : (BLK == 3) ? (unsigned)blk_off2<T>(n, a, a.in_row) : (unsigned)n * irow;
      v[m] = pass_in(ldg_pred(ip + off, valid), CA());
    }
  }
  PassIdx<N, TX, R1, CA> idx{c};
  // single exchange buffer: two barriers per exchange (scatter | gather | next scatter); twiddles straight from the
  // (L1-resident) global table -- a per-block shared copy does not pay off for one tile per block
  fft_run_sb<CA, N, E, DIR, 1, 1>(v, t, sm, TwGlobal<CA>{twp}, idx);
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int n = t + Tn * m;
    if (!PIN) {
      const bool hi = n >= a.hi0;
      const bool ok = valid && (n < a.lo || hi);
      const unsigned r = (unsigned)(n - (hi ? a.shift : 0));
      const unsigned off = (BLK == 2) ? (unsigned)blk_off((int)r, a.blk_rows, a.blk_stride, a.blk_magic, a.out_row)
                         : (BLK == 4) ? (unsigned)blk_off2<T>((int)r, a, a.out_row) : r * orow;
      stg_pred(op + off, pass_out(v[m]), ok);
    } else {
      const unsigned off = (BLK == 2) ? (unsigned)blk_off(n, a.blk_rows, a.blk_stride, a.blk_magic, a.out_row)
                         : (BLK == 4) ? (unsigned)blk_off2<T>(n, a, a.out_row) : (unsigned)n * orow;
      stg_pred(op + off, pass_out(v[m]), valid);
    }
  }
}

// ---- cross-rank flags of the slab exchange ----------------------------------------------------------------------
// A rank announces "my piece for slot s has landed in your receive buffer" by storing the exchange's epoch into word
// [s][sender] of the RECEIVER's flag array (peer memory mapped with CUDA IPC); the receiver's compute stream spins on its own
// array before it reads the buffer.  k_flag_set runs on the copy stream right behind the copy-engine push (stream order: the
// push has completed), k_flag_wait on the compute stream.  Replaces the two 1-element ncclAllReduce barriers per exchange of
// round 1 (40 tiny collectives per step).  A wait that sees nothing for MHDF_FLAG_TIMEOUT_NS gives up and reports through
// `err` (host-mapped): a lost peer becomes an error code, not a hung GPU.
#ifndef MHDF_FLAG_TIMEOUT_NS
#define MHDF_FLAG_TIMEOUT_NS 20000000000ULL
#endif
#ifdef MHDF_CPU_EMU
inline void st_flag(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned ld_flag(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#else
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_flag(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
static __global__ void __launch_bounds__(32) k_flag_set(unsigned* flag, unsigned v) {
  if (threadIdx.x == 0) { __threadfence_system(); st_flag(flag, v); }
}
// thread q < n (q != skip) waits until flags[q] has reached epoch v (wrap-safe comparison)
static __global__ void __launch_bounds__(32) k_flag_wait(const unsigned* flags, unsigned v, int n, int skip, int* err) {
  const int q = threadIdx.x;
  if (q < n && q != skip) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_flag(flags + q) - v) < 0) {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > MHDF_FLAG_TIMEOUT_NS) { *((volatile int*)err) = 1 + q; break; }
    }
  }
  __syncwarp();
  __threadfence_system();
}
#endif

// ------------------------------------------------------------------------------------------------
// x pass: rows are contiguous.  Real rows of length N are handled as M = N/2 complex points.
// ------------------------------------------------------------------------------------------------
// One pad slot per 16 elements: 16 consecutive elements (gathers, natural-order accesses, mirrored reads) and the
// first radix-8/16 scatter are bank-conflict free; offsets r*NS still fold into immediate operands.
template <int M, int R1> struct RowIdx {
  static constexpr int SIZE = M + M / 8 + 1;   // large enough for both maps
  __device__ __forceinline__ int operator()(int n) const { return n + (n >> 4); }
};
// Map of the second and later exchanges: 8 pad slots per 64 elements.  Their scatter writes 64*(t>>3) + (t&7) + 8r'
// (radix 8 after a radix-8 step): threads t and t+8 land in different bank halves, 16 consecutive elements stay
// conflict free, and offsets r*NS still fold into immediates.
template <int M> struct RowIdx2 {
  __device__ __forceinline__ int operator()(int n) const { return n + ((n >> 6) << 3); }
};

// Shared-memory context of one row: two alternating exchange buffers.
template <typename C> struct RowSmem {
  C* a;
  C* b;
  __device__ __forceinline__ void swap() { C* t = a; a = b; b = t; }
};

// c2r of one row: X[k], k < Kx (others zero) -> v = z[n] = x[2n] + i x[2n+1], n = t + Tm*m,
// unnormalised inverse times `scale`.  Im X[0] is ignored like every c2r does.
// Twiddles of the x kernels come from the global table; they do not depend on the row or the field, so the compiler
// keeps them in registers across the 15 transforms of a row set (a shared-memory copy measured 20 % slower).
template <typename T, int N, int E> struct XTwSrc {
  using C = Cx<T>;
  static __device__ __forceinline__ TwGlobal<C> fft(const C* tw) { return TwGlobal<C>{tw}; }
  static __device__ __forceinline__ C wn(const C* tw, int, int k) { return __ldg(tw + (unsigned)k); }
};

// Pre-step of the half-length c2r and post-step of the half-length r2c for one mode k (uniform in k, including k = 0 and
// k = M/2).  x1 = X[k], x2 = X[M-k] (not conjugated), w = exp(-2 pi i k / N); z1 = Z[k], z2 = Z[M-k].
//   c2r:  Z[k] = (x1 + conj x2) + i (x1 - conj x2) exp(+2 pi i k / N)
//   r2c:  X[k] = ((z1 + conj z2) - i (z1 - conj z2) w) / 2
// The default build keeps the operation order that was measured on hardware; the packed-FP32 build (MHDF_F32X2) uses the
// forms whose sign flips and lane swaps fold into FADD2 / FFMA2 operand modifiers (6 instructions each).
template <typename C, typename T>
__device__ __forceinline__ C c2r_pre(C x1, C x2, C w, T scale) {
#ifdef MHDF_F32X2
  const C x2c = cconj(x2);
  const C s = cadd(x1, x2c), d = cmulc(csub(x1, x2c), w);
  return cscale(cadd(s, cmuli(d)), scale);
#else
  x2 = cconj(x2);
  w = cconj(w);                                               // exp(+2 pi i k / N)
  const C s = cadd(x1, x2), d = cmul(csub(x1, x2), w);        // Z = s + i d
  return cscale(cadd(s, cmuli(d)), scale);
#endif
}
template <typename C>
__device__ __forceinline__ C r2c_post(C z1, C z2, C w) {
  using T = typename RealOf<C>::type;
#ifdef MHDF_F32X2
  const C z2c = cconj(z2);
  const C t = cmul(csub(z1, z2c), w);
  return cscale(cadd(cadd(z1, z2c), cmulmi(t)), (T)0.5);
#else
  z2 = cconj(z2);
  const C ev = cscale(cadd(z1, z2), (T)0.5);
  const C od = cscale(cmulmi(csub(z1, z2)), (T)0.5);          // -i (z1 - z2) / 2
  return cadd(ev, cmul(od, w));
#endif
}

// dkx != nullptr: the row is differentiated along x on the way in -- X[k] -> i kx[k] X[k] with the kr table of the grid
// (the EMHD kernels get d/dx B_i and d/dx A_i from the rows of B_i and A_i themselves instead of from separately
// transformed derivative fields: 6 of the 24 inverse-transformed fields of the gradient form disappear)
// DX is a compile-time switch: with a run-time one the compiler if-converts the derivative arithmetic into EVERY transform of a
// rolled loop (measured: the EMHD x pass went from 32.6 to 37.7 ms per 512^3 step, profiles/r02_c7_bench_emhd512.json).
template <typename T, int N, int E, typename SYNC, bool DX = false>
__device__ __forceinline__ void row_c2r(Cx<T> (&v)[E], const Cx<T>* __restrict__ X, int Kx, T scale, int t,
                                        RowSmem<Cx<T>>& sm, const Cx<T>* __restrict__ tw, const T* __restrict__ dkx = nullptr) {
  using C = Cx<T>;
  constexpr int M = N / 2, Tm = M / E, R1 = imin(E, M);
  MHDF_KEEP_PTR(X);
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int k = t + Tm * m;
    const int k2 = M - k;
    C x1 = ldg_pred(X + (unsigned)k, k < Kx);
    C x2 = ldg_pred(X + (unsigned)(k2 < Kx ? k2 : 0), k2 < Kx);
    if constexpr (DX) {   // same expression as k_emhd_derive: i * (k_x * f)
      x1 = cmuli(cscale(x1, k < Kx ? __ldg(dkx + k) : (T)0));
      x2 = cmuli(cscale(x2, k2 < Kx ? __ldg(dkx + k2) : (T)0));
    }
    if (k == 0) x1.y = 0;
    v[m] = c2r_pre(x1, x2, XTwSrc<T, N, E>::wn(tw, m, k), scale);
  }
  constexpr int NEX = fft_num_steps(M, E) - 1;
  fft_run<C, M, E, +1, 2, 1, 0, SYNC>(v, t, sm.a, sm.b, XTwSrc<T, N, E>::fft(tw), RowIdx<M, R1>(), RowIdx2<M>());
  if (NEX & 1) sm.swap();
}

// r2c of one row: v = z[n] (n = t + Tm*m) -> X[k] for k < Kx written to Xout (unnormalised forward).
template <typename T, int N, int E, typename SYNC>
__device__ __forceinline__ void row_r2c(Cx<T> (&v)[E], Cx<T>* __restrict__ Xout, int Kx, int t,
                                        RowSmem<Cx<T>>& sm, const Cx<T>* __restrict__ tw) {
  using C = Cx<T>;
  constexpr int M = N / 2, Tm = M / E, R1 = imin(E, M);
  constexpr int NEX = fft_num_steps(M, E) - 1;
  fft_run<C, M, E, -1, 2, 1, 0, SYNC>(v, t, sm.a, sm.b, XTwSrc<T, N, E>::fft(tw), RowIdx<M, R1>(), RowIdx2<M>());
  if (NEX & 1) sm.swap();
  if constexpr (Tm <= 32) {
    // Z[M-k] sits in the registers of lane (Tm - t) of the same row (slot E-1-m), or in this thread's own slot E-m when
    // t == 0: two warp shuffles per element replace the shared-memory round trip (and its barrier)
    MHDF_KEEP_PTR(Xout);
    const int lane = threadIdx.x & 31;
    const int src = (lane & ~(Tm - 1)) | ((Tm - t) & (Tm - 1));
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int k = t + Tm * m;
      const C z1 = v[m];
      C z2 = mk<C>(__shfl_sync(0xffffffffu, v[E - 1 - m].x, src), __shfl_sync(0xffffffffu, v[E - 1 - m].y, src));
      if (t == 0) z2 = v[(E - m) % E];
      stg_pred(Xout + (unsigned)k, r2c_post(z1, z2, XTwSrc<T, N, E>::wn(tw, m, k)), k < Kx);
    }
  } else {
    // Rows of two warps (1024-point rows): Z[M-k] crosses warps and goes through shared memory.  Measured and rejected on B200
    // (profiles/r02_c10_time1024.log): an unpadded buffer with stores / loads predicated to the elements a stored column k < Kx
    // reads -- fewer shared-memory wavefronts, but 1.5 % slower per 1024^3 x pass than this branch-free form.
    RowIdx<M, R1> idx;
#pragma unroll
    for (int m = 0; m < E; ++m) sm.a[idx(t + Tm * m)] = v[m];
    SYNC::sync();
    MHDF_KEEP_PTR(Xout);
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int k = t + Tm * m;
      // columns k >= Kx are never stored; their arithmetic is harmless and keeps the code branch-free
      const C z1 = v[m];
      const C z2 = sm.a[idx((M - k) & (M - 1))];
      stg_pred(Xout + (unsigned)k, r2c_post(z1, z2, XTwSrc<T, N, E>::wn(tw, m, k)), k < Kx);
    }
    sm.swap();
  }
}

// rows owned by <= 32 threads: warp barrier; otherwise the block must hold exactly one row
template <bool WARP, int RB> struct XSync { using type = SyncWarp; };
template <> struct XSync<false, 1> { using type = SyncBlock; };

// ---- physics functors: which products are formed between the inverse and forward x passes -------
// Reduction slots written by the fused x kernel (doubles): see XRed.
struct XRed {
  double sumsq[6];   // sum f^2 per input field (ux,uy,uz,bx,by,bz | EMHD: Ax,Ay,Az,bx,by,bz)
  double cross;      // sum u.b
  double nd;         // sum |u_i^2 f_i| over i = x, y, z (NDForceDriving!: the normalisation of the negative-damping force)
  unsigned long long maxsq[6]; // max f^2 per field as the bit pattern of a non-negative T (orders like an unsigned integer):
                               // Float32 in the low word, Float64 in the whole word -- getCFL!'s maximum in the problem's precision
};

template <typename T>
struct XArgs {
  const Cx<T>* in;     // [nin][rows][Kxp]
  Cx<T>* out;          // [nout][rows][Kxp]
  const Cx<T>* tw;     // length nx
  T* real_io;          // EMHD: stale real b [3][rows][nx] (read, then overwritten with fresh b); plain: real rows
  long long in_field, out_field, real_field;
  long long rows;      // ny*nz
  int Kx, Kxp;
  T scale;             // 1/(nx ny nz)
  XRed* red;           // may be null
  // volume penalisation (VP kernels only): real fields [1 + NF][rows][nx] = chi, then U0x,U0y,U0z[,B0x,B0y,B0z]
  const T* vp;
  long long vp_field;
  T vp_eta;            // eta = clock.dt * 13/7 (VPSolver.jl:23,45)
  const T* kxv;        // [Kx] kr of the grid (EMHD kernels: x derivatives are formed inside the row transform)
};

template <typename T> __device__ __forceinline__ void warp_red_sum(double& x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
}
__device__ __forceinline__ float max2(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max2(double a, double b) { return fmax(a, b); }
template <typename TM> __device__ __forceinline__ void warp_red_max(TM& x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = max2(x, __shfl_xor_sync(0xffffffffu, x, o));
}
__device__ __forceinline__ void atomic_max_bits(unsigned long long* p, float x) { atomicMax(reinterpret_cast<unsigned*>(p), __float_as_uint(x)); }   // little endian: low word
__device__ __forceinline__ void atomic_max_bits(unsigned long long* p, double x) { atomicMax(p, (unsigned long long)__double_as_longlong(x)); }

// Block reduction of NS sums + NM maxima followed by one atomic per quantity per block.
template <int NS, int NM, typename TM>
__device__ __forceinline__ void block_reduce_commit(double (&s)[NS], TM (&mx)[NM], double* gs, unsigned long long* gm) {
  __shared__ double sh_s[32][NS];
  __shared__ TM sh_m[32][NM];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NS; ++i) warp_red_sum<double>(s[i]);
#pragma unroll
  for (int i = 0; i < NM; ++i) warp_red_max(mx[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) sh_s[wid][i] = s[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) sh_m[wid][i] = mx[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double x = lane < nw ? sh_s[lane][i] : 0.0;
      warp_red_sum<double>(x);
      if (lane == 0) atomicAdd(&gs[i], x);
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      TM x = lane < nw ? sh_m[lane][i] : (TM)0;
      warp_red_max(x);
      if (lane == 0) atomic_max_bits(&gm[i], x);
    }
  }
}

enum { PHYS_HD = 0, PHYS_MHD = 1, PHYS_EMHD = 2 };

// Fused x pass:  c2r(NIN fields) -> pointwise products -> r2c(NOUT fields), one block = RB rows,
// grid-stride over row sets.  Real-space fields live only in registers.
//   HD  : in u(3)              out T_ij = -u_i u_j (xx,xy,xz,yy,yz,zz)
//   MHD : in u(3), b(3)        out T_ij = b_i b_j - u_i u_j (6), E = u x b (3)
//         (reference: MHDSolver.jl:73 and :150; HDSolver.jl:62)
//   EMHD: in A(3), d_{y,z}B(6), d_{y,z}A(6), B(3) spectral + stale b (3, real)   out G_i (3), fresh b -> real_io
//         G_i = sum_j A_j d_j B_i - b^stale_j d_j A_i   (reference: MHDSolver.jl:241-266, 323-325)
// RED: accumulate the per-field sum f^2 / max f^2 / sum u.b reductions (only the launch whose real-space fields
// the reference's stale `vars` correspond to needs them; the other stages skip the work and the registers)
// Register budget: measured on B200 (256^3 / 1024-wide rows), forcing more resident blocks (5, 6, 8 per SM) or 4
// elements per thread was 2-23 % slower than letting the kernel keep the whole row set in 255 registers.
//   VP  : HD / MHD with the volume-penalisation method: additionally out V_j = chi/eta (f_j - W_j), W = U0 (and B0 for the
//         magnetic field), j = x,y,z, appended after the tensor / E fields (reference: VPSolver.jl:21-59); its own
//         instantiation, the plain kernels are unchanged
//   ND  : MHD with NDForceDriving! (pgen/NegativeDamping.jl:23-45): additionally out F_i = f_i u_i (i = x, y, z) after the E fields and
//         the reduction sum |u_i^2 f_i| that normalises the force; VP = 2 selects it (VP = 1: volume penalisation)
template <typename T, int N, int E, int RB, int PHYS, bool RED, int VP = 0>
__global__ void __launch_bounds__((N / 2 / E) * RB) k_xfused(XArgs<T> a) {
  using C = Cx<T>;
  constexpr int M = N / 2, Tm = M / E, R1 = imin(E, M);
  // a row owned by at most one warp synchronises with __syncwarp only: rows are fully decoupled
  using SYNC = typename XSync<(Tm <= 32), RB>::type;
  MHDF_DYN_SMEM(unsigned char, smem_raw);
  const int r = threadIdx.x / Tm;
  const int t = threadIdx.x % Tm;
  constexpr int RS = RowIdx<M, R1>::SIZE;
  RowSmem<C> sm;
  sm.a = reinterpret_cast<C*>(smem_raw) + (size_t)(2 * r) * RS;
  sm.b = sm.a + RS;
  const C* twt = a.tw;
  MHDF_KEEP_PTR(twt);

  double rs[8];
  T rm[6];
#pragma unroll
  for (int i = 0; i < 8; ++i) rs[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) rm[i] = 0.f;

  const long long nsets = a.rows / RB;
  constexpr int NINF = (PHYS == PHYS_MHD) ? 6 : (PHYS == PHYS_HD ? 3 : 18);
  for (long long set = blockIdx.x; set < nsets; set += gridDim.x) {
    const long long row = set * RB + r;
    const C* in = a.in + row * a.Kxp;
    C* out = a.out + row * a.Kxp;
    {
      // pull the next row set of this block towards L2 while this one is being transformed
      const long long nset = set + gridDim.x;
      if (nset < nsets) {
        const char* nb = reinterpret_cast<const char*>(a.in + (nset * RB + r) * a.Kxp);
        const int bytes = a.Kx * (int)sizeof(C);
        for (int f = 0; f < NINF; ++f)
          for (int o = t * 128; o < bytes; o += Tm * 128)
#ifndef MHDF_CPU_EMU
            asm volatile("prefetch.global.L2 [%0];" :: "l"(nb + (long long)f * a.in_field * (long long)sizeof(C) + o));
#else
            (void)nb;
#endif
      }
    }
    if constexpr (PHYS == PHYS_HD || PHYS == PHYS_MHD) {
      constexpr int NF = (PHYS == PHYS_MHD) ? 6 : 3;
      C f[NF][E];
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        row_c2r<T, N, E, SYNC>(f[i], in + i * a.in_field, a.Kx, a.scale, t, sm, twt);
        if constexpr (RED) {
          T s = 0;
          T mx = 0;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const C sq = lmul(f[i][m], f[i][m]);
            s += sq.x + sq.y;
            mx = max2(mx, max2(sq.x, sq.y));
          }
          rs[i] += (double)s;
          rm[i] = max2(rm[i], mx);
        }
      }
      if constexpr (PHYS == PHYS_MHD && RED) {
        T s = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int m = 0; m < E; ++m) { const C ub = lmul(f[i][m], f[i + 3][m]); s += ub.x + ub.y; }
        rs[6] += (double)s;
      }
      // symmetric tensor (xx, xy, xz, yy, yz, zz)
      int p = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = i; j < 3; ++j) {
          C v[E];
#pragma unroll
          for (int m = 0; m < E; ++m) {
            if constexpr (PHYS == PHYS_MHD)
              v[m] = lmulsub(f[3 + i][m], f[3 + j][m], f[i][m], f[j][m]);
            else
              v[m] = lneg(lmul(f[i][m], f[j][m]));
          }
          row_r2c<T, N, E, SYNC>(v, out + p * a.out_field, a.Kx, t, sm, twt);
          ++p;
        }
      }
      if constexpr (PHYS == PHYS_MHD) {
        // E = u x b
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          constexpr int nx1[3] = {1, 2, 0}, nx2[3] = {2, 0, 1};
          const int j = nx1[i], k = nx2[i];
          C v[E];
#pragma unroll
          for (int m = 0; m < E; ++m)
            v[m] = lmulsub(f[j][m], f[3 + k][m], f[k][m], f[3 + j][m]);
          row_r2c<T, N, E, SYNC>(v, out + (6 + i) * a.out_field, a.Kx, t, sm, twt);
        }
      }
      if constexpr (VP == 2) {
        static_assert(PHYS == PHYS_MHD && RED, "NDForceDriving! acts on the MHD path; its normalisation needs the reduction epilogue");
        const T* frow = a.vp + row * (long long)N;
        T s = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          C v[E];
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const C fi = reinterpret_cast<const C*>(frow + i * a.vp_field)[t + Tm * m];
            v[m] = lmul(fi, f[i][m]);                                           // F_i = f_i u_i   (NegativeDamping.jl:38)
            const C w = lmul(lmul(f[i][m], f[i][m]), fi);                       // u_i^2 f_i        (:32-33)
            s += fabs(w.x) + fabs(w.y);
          }
          row_r2c<T, N, E, SYNC>(v, out + (9 + i) * a.out_field, a.Kx, t, sm, twt);
        }
        rs[7] += (double)s;
      }
      if constexpr (VP == 1) {
        constexpr int NT = (PHYS == PHYS_MHD) ? 9 : 6;
        const T* vrow = a.vp + row * (long long)N;
        C ce[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const C c = reinterpret_cast<const C*>(vrow)[t + Tm * m];
          ce[m] = mk<C>(c.x / a.vp_eta, c.y / a.vp_eta);                       // chi/eta
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          C v[E];
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const C w = reinterpret_cast<const C*>(vrow + (1 + i) * a.vp_field)[t + Tm * m];
            v[m] = lmul(ce[m], lsub(f[i][m], w));                               // chi/eta * (u_j - U_j)
          }
          row_r2c<T, N, E, SYNC>(v, out + (NT + i) * a.out_field, a.Kx, t, sm, twt);
        }
      }
    } else {
      // EMHD: field order in `in`: A(0..2), d_j B_i (j = y, z) at 3 + 2 i + (j - 1), d_j A_i (j = y, z) at 9 + 2 i + (j - 1),
      // B(15..17); the x derivatives come from the rows of B_i / A_i (row_c2r with the kr table)
      C A[3][E], bs[3][E];
      C* breal = reinterpret_cast<C*>(a.real_io + row * (long long)N);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        row_c2r<T, N, E, SYNC>(A[i], in + i * a.in_field, a.Kx, a.scale, t, sm, twt);
        T s = 0;
        T mx = 0;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if constexpr (RED) {
            const C sq = lmul(A[i][m], A[i][m]);
            s += sq.x + sq.y;
            mx = max2(mx, max2(sq.x, sq.y));
          }
          bs[i][m] = reinterpret_cast<const C*>(reinterpret_cast<const T*>(breal) + i * a.real_field)[t + Tm * m];
        }
        rs[i] += (double)s;
        rm[i] = max2(rm[i], mx);
      }
      // One rolled loop body for the three components: fully unrolled, the kernel carried 27 inlined row transforms = 194 KB of SASS
      // at 512-point rows and ran 60.4 ms per 512^3 step of x pass against 45.3 ms rolled (round 2, profiles/r02_c1_emhd.log) -- an
      // instruction-cache cliff.  Only the field offsets depend on i; A[j], bs[j] keep their compile-time indices.
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        C acc[E];
#pragma unroll
        for (int m = 0; m < E; ++m) acc[m] = mk<C>(0, 0);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          C g[E];
          if (j == 0) row_c2r<T, N, E, SYNC, true>(g, in + (15 + i) * a.in_field, a.Kx, a.scale, t, sm, twt, a.kxv);
          else row_c2r<T, N, E, SYNC>(g, in + (3 + 2 * i + (j - 1)) * a.in_field, a.Kx, a.scale, t, sm, twt);
#pragma unroll
          for (int m = 0; m < E; ++m) acc[m] = lfma(A[j][m], g[m], acc[m]);
          if (j == 0) row_c2r<T, N, E, SYNC, true>(g, in + i * a.in_field, a.Kx, a.scale, t, sm, twt, a.kxv);
          else row_c2r<T, N, E, SYNC>(g, in + (9 + 2 * i + (j - 1)) * a.in_field, a.Kx, a.scale, t, sm, twt);
#pragma unroll
          for (int m = 0; m < E; ++m) acc[m] = lfma(lneg(bs[j][m]), g[m], acc[m]);
        }
        row_r2c<T, N, E, SYNC>(acc, out + i * a.out_field, a.Kx, t, sm, twt);
      }
      // refresh the real-space b (vars.b*) from the current stage input
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        C g[E];
        row_c2r<T, N, E, SYNC>(g, in + (15 + i) * a.in_field, a.Kx, a.scale, t, sm, twt);
        T s = 0;
        T mx = 0;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if constexpr (RED) {
            const C sq = lmul(g[m], g[m]);
            s += sq.x + sq.y;
            mx = max2(mx, max2(sq.x, sq.y));
          }
          reinterpret_cast<C*>(reinterpret_cast<T*>(breal) + i * a.real_field)[t + Tm * m] = g[m];
        }
        rs[3 + i] += (double)s;
        rm[3 + i] = max2(rm[3 + i], mx);
      }
    }
  }
  if constexpr (RED) {
    if (a.red != nullptr) block_reduce_commit<8, 6>(rs, rm, a.red->sumsq, a.red->maxsq);
  }
}

// EMHD fused x pass, second form (opt-in: MHDF_EMHD2=1).  Same arithmetic in the same order as the EMHD branch of k_xfused --
// results are bit-identical -- but the six multiplier fields (A_j and the stale b_j in real space) wait in thread-private
// shared-memory slots instead of 96 registers, so every loop can stay rolled: 5 inlined row transforms instead of 13 (27
// before round 1's last change), a 168-register budget (6 instead of 4 resident blocks per SM) and 58-71 KB of SASS instead of
// 103-118 KB (180-203 KB).  Slot of (field f, element m) of a thread: mult[(f * E + m) * Tm] with mult already offset by the
// thread's (row, t): consecutive threads hit consecutive banks; only the owning thread ever touches a slot, no barrier.
#ifndef MHDF_EMHD2_MINB
#define MHDF_EMHD2_MINB 6   // 168 registers: 16 bytes of spills at 512-point rows (128 registers: 108-224 bytes); 6 blocks x 33 KB of shared memory fill an SM
#endif
template <typename T, int N, int E, int RB, bool RED>
__global__ void __launch_bounds__((N / 2 / E) * RB, (N / 2 / E) * RB * MHDF_EMHD2_MINB <= 1024 ? MHDF_EMHD2_MINB : 1) k_xfused_emhd2(XArgs<T> a) {
  using C = Cx<T>;
  constexpr int M = N / 2, Tm = M / E, R1 = imin(E, M);
  using SYNC = typename XSync<(Tm <= 32), RB>::type;
  MHDF_DYN_SMEM(unsigned char, smem_raw);
  const int r = threadIdx.x / Tm;
  const int t = threadIdx.x % Tm;
  constexpr int RS = RowIdx<M, R1>::SIZE;
  RowSmem<C> sm;
  sm.a = reinterpret_cast<C*>(smem_raw) + (size_t)(2 * r) * RS;
  sm.b = sm.a + RS;
  C* mult = reinterpret_cast<C*>(smem_raw) + (size_t)2 * RB * RS + (size_t)r * 6 * M + t;
  const C* twt = a.tw;
  MHDF_KEEP_PTR(twt);
  double rs[8];
  T rm[6];
#pragma unroll
  for (int i = 0; i < 8; ++i) rs[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) rm[i] = 0.f;
  // reductions with a run-time slot: predicated adds keep rs / rm in registers
  auto red_add = [&](int slot, T s, T mx) {
#pragma unroll
    for (int q = 0; q < 6; ++q)
      if (q == slot) { rs[q] += (double)s; rm[q] = max2(rm[q], mx); }
  };
  const long long nsets = a.rows / RB;
  for (long long set = blockIdx.x; set < nsets; set += gridDim.x) {
    const long long row = set * RB + r;
    const C* in = a.in + row * a.Kxp;
    C* out = a.out + row * a.Kxp;
    C* breal = reinterpret_cast<C*>(a.real_io + row * (long long)N);
    {
      const long long nset = set + gridDim.x;
      if (nset < nsets) {
        const char* nb = reinterpret_cast<const char*>(a.in + (nset * RB + r) * a.Kxp);
        const int bytes = a.Kx * (int)sizeof(C);
        for (int f = 0; f < 18; ++f)
          for (int o = t * 128; o < bytes; o += Tm * 128)
#ifndef MHDF_CPU_EMU
            asm volatile("prefetch.global.L2 [%0];" :: "l"(nb + (long long)f * a.in_field * (long long)sizeof(C) + o));
#else
            (void)nb;
#endif
      }
    }
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {   // A_i = c2r(A^_i) and the stale b_i -> multiplier slots
      C g[E];
      row_c2r<T, N, E, SYNC>(g, in + i * a.in_field, a.Kx, a.scale, t, sm, twt);
      T s = 0;
      T mx = 0;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        if constexpr (RED) {
          const C sq = lmul(g[m], g[m]);
          s += sq.x + sq.y;
          mx = max2(mx, max2(sq.x, sq.y));
        }
        mult[(i * E + m) * Tm] = g[m];
        mult[((3 + i) * E + m) * Tm] = reinterpret_cast<const C*>(reinterpret_cast<const T*>(breal) + i * a.real_field)[t + Tm * m];
      }
      if constexpr (RED) red_add(i, s, mx);
    }
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {   // G_i = sum_j A_j d_j B_i - b^stale_j d_j A_i
      C acc[E];
#pragma unroll
      for (int m = 0; m < E; ++m) acc[m] = mk<C>(0, 0);
      {   // j = x: the rows of B_i / A_i themselves, differentiated on the way in (its own two inlined transforms)
        C g[E];
        row_c2r<T, N, E, SYNC, true>(g, in + (15 + i) * a.in_field, a.Kx, a.scale, t, sm, twt, a.kxv);
#pragma unroll
        for (int m = 0; m < E; ++m) acc[m] = lfma(mult[m * Tm], g[m], acc[m]);
        row_c2r<T, N, E, SYNC, true>(g, in + i * a.in_field, a.Kx, a.scale, t, sm, twt, a.kxv);
#pragma unroll
        for (int m = 0; m < E; ++m) acc[m] = lfma(lneg(mult[(3 * E + m) * Tm]), g[m], acc[m]);
      }
#pragma unroll 1
      for (int j = 1; j < 3; ++j) {
        C g[E];
        row_c2r<T, N, E, SYNC>(g, in + (3 + 2 * i + (j - 1)) * a.in_field, a.Kx, a.scale, t, sm, twt);
#pragma unroll
        for (int m = 0; m < E; ++m) acc[m] = lfma(mult[(j * E + m) * Tm], g[m], acc[m]);
        row_c2r<T, N, E, SYNC>(g, in + (9 + 2 * i + (j - 1)) * a.in_field, a.Kx, a.scale, t, sm, twt);
#pragma unroll
        for (int m = 0; m < E; ++m) acc[m] = lfma(lneg(mult[((3 + j) * E + m) * Tm]), g[m], acc[m]);
      }
      row_r2c<T, N, E, SYNC>(acc, out + i * a.out_field, a.Kx, t, sm, twt);
    }
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {   // refresh the real-space b (vars.b*) from the current stage input
      C g[E];
      row_c2r<T, N, E, SYNC>(g, in + (15 + i) * a.in_field, a.Kx, a.scale, t, sm, twt);
      T s = 0;
      T mx = 0;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        if constexpr (RED) {
          const C sq = lmul(g[m], g[m]);
          s += sq.x + sq.y;
          mx = max2(mx, max2(sq.x, sq.y));
        }
        reinterpret_cast<C*>(reinterpret_cast<T*>(breal) + i * a.real_field)[t + Tm * m] = g[m];
      }
      if constexpr (RED) red_add(3 + i, s, mx);
    }
  }
  if constexpr (RED) {
    if (a.red != nullptr) block_reduce_commit<8, 6>(rs, rm, a.red->sumsq, a.red->maxsq);
  }
}

// Plain x passes for the API boundary (set_real / get_real): real rows <-> spectral rows.
template <typename T, int N, int E, int RB, int DIR>
__global__ void __launch_bounds__((N / 2 / E) * RB) k_xplain(XArgs<T> a) {
  using C = Cx<T>;
  constexpr int M = N / 2, Tm = M / E, R1 = imin(E, M);
  using SYNC = typename XSync<(Tm <= 32), RB>::type;
  MHDF_DYN_SMEM(unsigned char, smem_raw);
  const int r = threadIdx.x / Tm;
  const int t = threadIdx.x % Tm;
  constexpr int RS = RowIdx<M, R1>::SIZE;
  RowSmem<C> sm;
  sm.a = reinterpret_cast<C*>(smem_raw) + (size_t)(2 * r) * RS;
  sm.b = sm.a + RS;
  const C* twt = a.tw;
  MHDF_KEEP_PTR(twt);
  double rs[1] = {0.0};
  T rm[1] = {(T)0};
  const long long nsets = a.rows / RB;
  for (long long set = blockIdx.x; set < nsets; set += gridDim.x) {
    const long long row = set * RB + r;
    C* re = reinterpret_cast<C*>(a.real_io + row * (long long)N);
    C v[E];
    if constexpr (DIR < 0) {   // real -> spectral
      T s = 0;
      T mx = 0;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        v[m] = re[t + Tm * m];
        const T x2 = v[m].x * v[m].x, y2 = v[m].y * v[m].y;
        s += x2 + y2;
        mx = max2(mx, max2(x2, y2));
      }
      rs[0] += (double)s;
      rm[0] = max2(rm[0], mx);
      row_r2c<T, N, E, SYNC>(v, a.out + row * a.Kxp, a.Kx, t, sm, twt);
    } else {                   // spectral -> real
      row_c2r<T, N, E, SYNC>(v, a.in + row * a.Kxp, a.Kx, a.scale, t, sm, twt);
      T s = 0;
      T mx = 0;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        re[t + Tm * m] = v[m];
        const T x2 = v[m].x * v[m].x, y2 = v[m].y * v[m].y;
        s += x2 + y2;
        mx = max2(mx, max2(x2, y2));
      }
      rs[0] += (double)s;
      rm[0] = max2(rm[0], mx);
    }
  }
  if (a.red != nullptr) block_reduce_commit<1, 1>(rs, rm, a.red->sumsq, a.red->maxsq);
}

// ------------------------------------------------------------------------------------------------
// Spectral kernels on the compact state.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct SpecGeom {
  int Kx, Kxp;
  Band by, bz;   // global retained bands
  int Kyl;       // local compact ky rows (slab of this rank; == by.count() on one GPU; padded rows hold zeros)
  int ky0;       // global compact row of local row 0
  int F;         // fields in the state
  const T* kx;   // [Kx]   kr  (reference: grid.kr)
  const T* ky;   // [Kyl]  l on the local retained rows
  const T* kz;   // [Kz]   m on the retained rows
  long long field;   // elements per local compact field = Kxp*Kyl*Kz
  // slab runs: kr = 0 plane of the stage input gathered from all ranks, [rank][F][Kz][Kyl]; null on one GPU
  const Cx<T>* mirror;
};

// element (kx, jc, kc) of a compact field; û^sym on the kr = 0 plane:
// rfft(irfft(û)) = (û(0,ky,kz) + conj û(0,-ky,-kz)) / 2, a missing (dealiased) mirror counts as 0
// (reference: the diffusion operand of MHDSolver.jl:91-94,166-168 / HDSolver.jl:82-85).
template <typename T>
__device__ __forceinline__ Cx<T> load_sym(const Cx<T>* __restrict__ S, int fi, const SpecGeom<T>& g, int ix, int jc, int kc) {
  using C = Cx<T>;
  const C* f = S + fi * g.field;
  C v = f[((long long)kc * g.Kyl + jc) * g.Kxp + ix];
  if (ix == 0) {
    const int jm = g.by.row_of_wave(-g.by.wave(g.ky0 + jc));   // global compact row of the mirror mode
    const int km = g.bz.row_of_wave(-g.bz.wave(kc));
    C w = mk<C>(0, 0);
    if (jm >= 0 && km >= 0) {
      if (g.mirror != nullptr) {
        const int q = jm / g.Kyl, jl = jm - q * g.Kyl;
        w = g.mirror[(((long long)q * g.F + fi) * g.bz.count() + km) * g.Kyl + jl];
      } else {
        w = f[((long long)km * g.Kyl + jm) * g.Kxp];
      }
    }
    v = mk<C>((T)0.5 * (v.x + w.x), (T)0.5 * (v.y - w.y));
  }
  return v;
}

// kr = 0 plane of a compact state, [F][Kz][Kyl] (input of the mirror all-gather)
template <typename T>
__global__ void __launch_bounds__(256) k_plane(SpecGeom<T> g, const Cx<T>* __restrict__ S, Cx<T>* __restrict__ out) {
  const long long total = (long long)g.F * g.bz.count() * g.Kyl;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long fk = e / g.Kyl;           // f * Kz + kc
    const int jc = (int)(e - fk * g.Kyl);
    const long long f = fk / g.bz.count(), kc = fk - f * g.bz.count();
    out[e] = S[f * g.field + (kc * g.Kyl + jc) * g.Kxp];
  }
}

// ---- A99 random solenoidal driving (Alvelius 1999): the reference's `calcF!` = A99ForceDriving! ------------------------
// Counter-based random numbers: Philox4x32-10 (Salmon et al. 2011), key = seed, counter = (global mode index in the
// reference's (nkr, ny, nz) array, RHS-evaluation number): the stream depends neither on the compact layout nor on the
// slab decomposition, and the CPU test-suite regenerates it bit for bit in NumPy.
struct Philox4 { unsigned v[4]; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ULL * c.v[0], p1 = 0xCD9E8D57ULL * c.v[2];
    Philox4 n;
    n.v[0] = (unsigned)(p1 >> 32) ^ c.v[1] ^ k0;
    n.v[1] = (unsigned)p1;
    n.v[2] = (unsigned)(p0 >> 32) ^ c.v[3] ^ k1;
    n.v[3] = (unsigned)p0;
    c = n;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
__host__ __device__ __forceinline__ float u01(unsigned a, unsigned, float) { return (float)(a >> 8) * 5.9604644775390625e-8f; }   // 24 bits
__host__ __device__ __forceinline__ double u01(unsigned a, unsigned b, double) {                                                    // 53 bits
  return (double)((((unsigned long long)a << 32) | b) >> 11) * 1.1102230246251565404e-16;
}
enum { A99_OFF = 0, A99_HOST = 1, A99_GPU = 2 };
template <typename T>
struct A99Args {
  int variant;       // A99_HOST: top-level A99ForceDriving! (pgen/A99ForceDriving.jl:33-60, tables of SetUpFk :93-127)
                     // A99_GPU:  module A99GPU (pgen/A99ForceDriving_GPU.jl:49-130)
  int nkr;           // nx/2 + 1 (mode numbering)
  T amp;             // usr_vars.A times the A inside Fk
  T kf, sig2, b, itanh;   // itanh = 1 / tanh(b pi/2)
  unsigned seed_lo, seed_hi;
  unsigned call_lo, call_hi;   // RHS-evaluation number
};
// four uniforms in [0,1) of mode `mode` for this RHS evaluation
template <typename T>
__host__ __device__ __forceinline__ void a99_uniforms(const A99Args<T>& q, unsigned long long mode, T (&r)[4]) {
  Philox4 c;
  c.v[0] = (unsigned)mode; c.v[1] = (unsigned)(mode >> 32); c.v[2] = q.call_lo; c.v[3] = q.call_hi << 1;
  const Philox4 a = philox4x32_10(c, q.seed_lo, q.seed_hi);
  if constexpr (sizeof(T) == 4) {
    r[0] = u01(a.v[0], 0u, T()); r[1] = u01(a.v[1], 0u, T()); r[2] = u01(a.v[2], 0u, T()); r[3] = u01(a.v[3], 0u, T());
  } else {
    c.v[3] |= 1u;
    const Philox4 b = philox4x32_10(c, q.seed_lo, q.seed_hi);
    r[0] = u01(a.v[0], a.v[1], T()); r[1] = u01(a.v[2], a.v[3], T()); r[2] = u01(b.v[0], b.v[1], T()); r[3] = u01(b.v[2], b.v[3], T());
  }
}
// c * x + y per component as ONE fused multiply-add (the same rounding whatever the compiler would contract)
__device__ __forceinline__ float2 axpy(float c, float2 x, float2 y) { return mk<float2>(fmaf(c, x.x, y.x), fmaf(c, x.y, y.y)); }
__device__ __forceinline__ double2 axpy(double c, double2 x, double2 y) { return mk<double2>(fma(c, x.x, y.x), fma(c, x.y, y.y)); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }     // a product that is never contracted into an FMA
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ void sincos2pi(float r, float& s, float& c) { sincospif(2.f * r, &s, &c); }
__device__ __forceinline__ void sincos2pi(double r, double& s, double& c) { sincospi(2.0 * r, &s, &c); }
// The forcing of one mode: f[0..2] += amp Fk (e^{i th1} g_i e1 + e^{i th2} g_j e2), Fk = sqrt(exp(-(k-kf)^2/sig2)/2pi)/k.
//   A99_HOST: e1 = (ky,-kx,0)/kp, e2 = (kx kz, ky kz, -kp^2)/(kp k), kp^2 = kx^2+ky^2, 0/0 -> 0, Fk = 0 on the kr = 0 plane
//             (:112-125).  SetUpFk builds e1x, e1y as (nkr, nl, 1) arrays and copyto!()s them into (nkr, nl, nm) tables
//             (:121-122): only the first z plane (kz index 0) of the e1 tables is ever filled -- `e1_plane` says whether
//             this mode lies on it.  Phi = pi rand() is drawn into the COMPLEX scratch array nonlinh1 (:46), so g_i = -tanh(b(Phi-pi/2))/
//             tanh(b pi/2) and g_j = sqrt(1-g_i^2) are complex numbers there -- restated literally.
//   A99_GPU:  e1 = (kz,0,-kx)/kp, e2 = (kx ky, -kp^2, kz ky)/(kp k), kp^2 = kx^2+kz^2, real Phi, |g_i| clipped to 1
//             (:92-113); afterwards Im N_u = 0 on the kr = 0 plane (:115-121) -- applied by the caller.
template <typename T>
__device__ __forceinline__ void a99_force(const A99Args<T>& q, unsigned long long mode, int ix, bool e1_plane, T kx, T ky, T kz, Cx<T> (&f)[3]) {
  using C = Cx<T>;
  const T PI = (T)3.14159265358979323846;
  f[0] = f[1] = f[2] = mk<C>(0, 0);
  if (q.variant == A99_HOST && ix == 0) return;
  const T k2 = kx * kx + ky * ky + kz * kz;
  if (!(k2 > (T)0)) return;
  const T k = sqrt(k2), ik = (T)1 / k, d = k - q.kf;
  const T Fk = q.amp * sqrt(exp(-(d * d) / q.sig2) / (T)2 / PI) * ik;
  T r[4];
  a99_uniforms<T>(q, mode, r);
  T s1, c1, s2, c2;
  sincos2pi(r[0], s1, c1);
  sincos2pi(r[3], s2, c2);
  T e1[3], e2[3];
  C gi, gj;
  if (q.variant == A99_HOST) {
    const T kp = sqrt(kx * kx + ky * ky);
    const bool ok = kp > (T)0;
    e1[0] = (ok && e1_plane) ? ky / kp : (T)0; e1[1] = (ok && e1_plane) ? -kx / kp : (T)0; e1[2] = (T)0;
    e2[0] = ok ? kx * kz / kp * ik : (T)0; e2[1] = ok ? ky * kz / kp * ik : (T)0; e2[2] = -kp * ik;
    // tanh(x + i y) by Kahan's formula (no cancellation next to the poles x = 0, y = pi/2 + n pi, where |g_i| reaches 1e3):
    //   t = tan y, beta = 1 + t^2, s = sinh x, rho = sqrt(1 + s^2):  tanh = (beta rho s + i t) / (1 + beta s^2)
    // x + i y = b (Phi - pi/2), Phi = pi (r1 + i r2); the products are rounded separately (no FMA contraction), like the
    // reference's broadcast `Phi *= pi; b*(Phi - pi/2)` -- next to a pole the result amplifies argument rounding 1e3-fold
    const T x = q.b * (mul_rn(PI, r[1]) - PI / (T)2), y = q.b * mul_rn(PI, r[2]);
    const T tn = tan(y), beta = (T)1 + tn * tn, sh = sinh(x), rho = sqrt((T)1 + sh * sh);
    const T den = (T)1 + beta * sh * sh;
    gi = mk<C>(-(beta * rho * sh / den) * q.itanh, -(tn / den) * q.itanh);
    // principal sqrt(1 - gi^2)
    const T u = (T)1 - (gi.x * gi.x - gi.y * gi.y), v = -(T)2 * gi.x * gi.y;
    const T m = sqrt(u * u + v * v);
    if (m == (T)0) gj = mk<C>(0, 0);
    else if (u >= (T)0) { const T s = sqrt((T)0.5 * (m + u)); gj = mk<C>(s, v / ((T)2 * s)); }
    else { const T s = sqrt((T)0.5 * (m - u)); gj = mk<C>(fabs(v) / ((T)2 * s), v < (T)0 ? -s : s); }
  } else {
    const T kp = sqrt(kx * kx + kz * kz);
    const bool ok = kp > (T)0;
    e1[0] = ok ? kz / kp : (T)0; e1[1] = (T)0; e1[2] = ok ? -kx / kp : (T)0;
    e2[0] = ok ? kx * ky / kp * ik : (T)0; e2[1] = -kp * ik; e2[2] = ok ? kz * ky / kp * ik : (T)0;
    T g = -tanh(q.b * (mul_rn(r[1], PI) - PI / (T)2)) * q.itanh;
    if (fabs(g) >= (T)1) g = g < (T)0 ? (T)-1 : (T)1;
    gi = mk<C>(g, 0);
    gj = mk<C>(sqrt((T)1 - g * g), 0);
  }
  const C p1 = mk<C>(c1 * gi.x - s1 * gi.y, c1 * gi.y + s1 * gi.x);   // e^{i th1} g_i
  const C p2 = mk<C>(c2 * gj.x - s2 * gj.y, c2 * gj.y + s2 * gj.x);   // e^{i th2} g_j
#pragma unroll
  for (int c = 0; c < 3; ++c) f[c] = mk<C>(Fk * (p1.x * e1[c] + p2.x * e2[c]), Fk * (p1.y * e1[c] + p2.y * e2[c]));
}

enum { STEP_CALCN = 0, STEP_RK4_1 = 1, STEP_RK4_2 = 2, STEP_RK4_3 = 3, STEP_RK4_4 = 4, STEP_LSRK = 5 };

template <typename T>
struct SpecArgs {
  SpecGeom<T> g;
  const Cx<T>* P;        // product spectra [nout][compact]
  const Cx<T>* Sin;      // stage input  [F][compact]
  const Cx<T>* Y;        // step start   [F][compact]   (RK4)
  Cx<T>* Sout;           // next stage input / LSRK: updated sol
  Cx<T>* A;              // RK4 accumulator / LSRK: S2 register
  Cx<T>* Nout;           // STEP_CALCN: RHS output
  T nu, eta;
  int n_nu;
  T ca, cs;              // RK4: A-weight (dt/6, dt/3), stage coefficient (dt/2, dt); LSRK: ca = A_i, cs = B_i
  T dt;
  int mode;
  int first;             // LSRK: stage 1 (S2 treated as zero)
  const Cx<T>* force;    // constant spectral forcing [F][compact] (calcF! hook), or null
  unsigned fmask;        // bit f set: field f is forced
  A99Args<T> a99;        // random driving (variant = A99_OFF: none)
  const double* nd_sum;  // NDForceDriving!: sum |u_i^2 f_i| of this evaluation (device, complete when the kernel starts)
  double nd_P;           // its P / dV
};
// Volume penalisation in the spectral kernel (VP instantiations): the penalisation spectra V^_j follow the tensor / E fields
// in P;  N_a += -sum_j (delta_aj - k_j k_a / k^2) V^_j  for the velocity and, in MHD, the same with the B0 set for the
// magnetic field (reference: VPSolver.jl:36, :56, called from HDSolver.jl:77-79, MHDSolver.jl:86-88, 161-163).

// RHS assembly + Runge-Kutta stage update, one thread per retained mode.
//   MHD/HD: N_a = D_a - k_a (k.D)/k^2 - nu k^2 u^sym_a [- nu k^(2 n_nu) u^sym_a],  D_j = sum_i i k_i T^_ij
//           N_{3+a} = i (k x E^)_a - eta k^2 b^sym_a
//   (reference: MHDSolver.jl:77-79,94,97-99,155,168; HDSolver.jl:68-70,85,88-90)
//   EMHD:   N_i = G^_i   (reference: MHDSolver.jl:253,264; no resistive term on this path)
#ifndef MHDF_SPEC_MINB
#define MHDF_SPEC_MINB 4   // <= 64 registers: a streaming kernel wants the occupancy, not the registers
#endif
// Mirror operand of the kr = 0 symmetrisation, resolved once per mode for all fields: base pointer of
// field 0 and the field stride, or null when the mode is off the plane / its mirror is dealiased.
template <typename T> struct SymSrc {
  const Cx<T>* base;
  long long stride;
  bool plane;
};
template <typename T>
__device__ __forceinline__ SymSrc<T> sym_src(const Cx<T>* __restrict__ S, const SpecGeom<T>& g, int ix, int jc, int kc) {
  SymSrc<T> r;
  r.base = nullptr; r.stride = 0; r.plane = (ix == 0);
  if (r.plane) {
    const int jm = g.by.row_of_wave(-g.by.wave(g.ky0 + jc));
    const int km = g.bz.row_of_wave(-g.bz.wave(kc));
    if (jm >= 0 && km >= 0) {
      if (g.mirror != nullptr) {
        const int q = jm / g.Kyl, jl = jm - q * g.Kyl;
        r.base = g.mirror + (((long long)q * g.F) * g.bz.count() + km) * g.Kyl + jl;
        r.stride = (long long)g.bz.count() * g.Kyl;
      } else {
        r.base = S + ((long long)km * g.Kyl + jm) * g.Kxp;
        r.stride = g.field;
      }
    }
  }
  return r;
}
template <typename T>
__device__ __forceinline__ Cx<T> sym_apply(const SymSrc<T>& r, int fi, Cx<T> v) {
  using C = Cx<T>;
  if (r.plane) {
    const C w = (r.base != nullptr) ? r.base[fi * r.stride] : mk<C>(0, 0);
    v = mk<C>((T)0.5 * (v.x + w.x), (T)0.5 * (v.y - w.y));
  }
  return v;
}

// RHS of one retained mode e = (ix, jc, kc): N[] and the stage input sin[] at that mode.  The mirror operand of the kr = 0
// symmetrisation is resolved once for all fields and the stage input is loaded once.
template <typename T, int PHYS, bool A99, int VP, typename IDX>
__device__ __forceinline__ void spec_rhs(const SpecArgs<T>& a, IDX e, int ix, int jc, int kc,
                                         Cx<T> (&N)[PHYS == PHYS_MHD ? 6 : 3], Cx<T> (&sin)[PHYS == PHYS_MHD ? 6 : 3]) {
  using C = Cx<T>;
  const SpecGeom<T>& g = a.g;
  if constexpr (PHYS == PHYS_EMHD) {
#pragma unroll
    for (int f = 0; f < 3; ++f) { N[f] = a.P[f * g.field + e]; sin[f] = a.Sin[f * g.field + e]; }
  } else {
    constexpr int F = (PHYS == PHYS_MHD) ? 6 : 3;
    const T kx = g.kx[ix], ky = g.ky[jc], kz = g.kz[kc];
    const T k2 = kx * kx + ky * ky + kz * kz;
    const T ik2 = (k2 > (T)0) ? (T)1 / k2 : (T)0;
    C Tt[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) Tt[p] = a.P[p * g.field + e];
    // D_j = i sum_i k_i T_ij   (xx,xy,xz,yy,yz,zz)
    C D[3];
    D[0] = mk<C>(kx * Tt[0].x + ky * Tt[1].x + kz * Tt[2].x, kx * Tt[0].y + ky * Tt[1].y + kz * Tt[2].y);
    D[1] = mk<C>(kx * Tt[1].x + ky * Tt[3].x + kz * Tt[4].x, kx * Tt[1].y + ky * Tt[3].y + kz * Tt[4].y);
    D[2] = mk<C>(kx * Tt[2].x + ky * Tt[4].x + kz * Tt[5].x, kx * Tt[2].y + ky * Tt[4].y + kz * Tt[5].y);
    const C kD = mk<C>((kx * D[0].x + ky * D[1].x + kz * D[2].x) * ik2, (kx * D[0].y + ky * D[1].y + kz * D[2].y) * ik2);
    const T kk[3] = {kx, ky, kz};
    const SymSrc<T> sy = sym_src<T>(a.Sin, g, ix, jc, kc);
#pragma unroll
    for (int f = 0; f < F; ++f) sin[f] = a.Sin[f * g.field + e];
    T hyper = (T)0;
    if (a.n_nu > 1) { hyper = (T)1; for (int q = 0; q < a.n_nu; ++q) hyper *= k2; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const C pr = mk<C>(D[c].x - kk[c] * kD.x, D[c].y - kk[c] * kD.y);   // still missing the factor i
      const C us = sym_apply<T>(sy, c, sin[c]);
      const T dc = -(a.nu * k2) - a.nu * hyper;
      N[c] = mk<C>(-pr.y + dc * us.x, pr.x + dc * us.y);
    }
    if constexpr (PHYS == PHYS_MHD) {
      C Ev[3];
#pragma unroll
      for (int p = 0; p < 3; ++p) Ev[p] = a.P[(6 + p) * g.field + e];
      C Cv[3];
      Cv[0] = mk<C>(ky * Ev[2].x - kz * Ev[1].x, ky * Ev[2].y - kz * Ev[1].y);
      Cv[1] = mk<C>(kz * Ev[0].x - kx * Ev[2].x, kz * Ev[0].y - kx * Ev[2].y);
      Cv[2] = mk<C>(kx * Ev[1].x - ky * Ev[0].x, kx * Ev[1].y - ky * Ev[0].y);
      const T dc = -(a.eta * k2);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const C bsym = sym_apply<T>(sy, 3 + c, sin[3 + c]);
        N[3 + c] = mk<C>(-Cv[c].y + dc * bsym.x, Cv[c].x + dc * bsym.y);
      }
    }
    if constexpr (VP == 1) {
      constexpr int NT = (PHYS == PHYS_MHD) ? 9 : 6, NG = (PHYS == PHYS_MHD) ? 2 : 1;
#pragma unroll
      for (int gp = 0; gp < NG; ++gp) {
        C V[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) V[j] = a.P[(NT + 3 * gp + j) * g.field + e];
        const C kV = mk<C>((kx * V[0].x + ky * V[1].x + kz * V[2].x) * ik2, (kx * V[0].y + ky * V[1].y + kz * V[2].y) * ik2);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          N[3 * gp + c].x -= V[c].x - kk[c] * kV.x;
          N[3 * gp + c].y -= V[c].y - kk[c] * kV.y;
        }
      }
    }
    if constexpr (PHYS == PHYS_MHD) {   // addforcing! after the advection (pgen.jl:159); HD / EMHD: no effect, like the reference
      if constexpr (VP == 2) {   // NDForceDriving!: N_ui += A F^_i, A = P / (sum |u_i^2 f_i| dV)   (NegativeDamping.jl:34-41)
        const T A = (T)(a.nd_P / *a.nd_sum);
#pragma unroll
        for (int c = 0; c < 3; ++c) { const C w = a.P[(9 + c) * g.field + e]; N[c].x += A * w.x; N[c].y += A * w.y; }
      }
      if (a.force != nullptr) {
#pragma unroll
        for (int f = 0; f < F; ++f)
          if ((a.fmask >> f) & 1u) { const C w = a.force[f * g.field + e]; N[f].x += w.x; N[f].y += w.y; }
      }
      if constexpr (A99) {   // its own instantiation: the undriven kernels keep their register budget and instruction stream
        const int jg = g.ky0 + jc;
        const unsigned iy = (unsigned)(jg < g.by.lo ? jg : jg + (g.by.hi0 - g.by.lo));
        const unsigned iz = (unsigned)(kc < g.bz.lo ? kc : kc + (g.bz.hi0 - g.bz.lo));
        const unsigned long long mode = (unsigned)ix + (unsigned long long)a.a99.nkr * (iy + (unsigned long long)g.by.n * iz);
        C fr[3];
        a99_force<T>(a.a99, mode, ix, iz == 0u, kx, ky, kz, fr);
#pragma unroll
        for (int c = 0; c < 3; ++c) { N[c].x += fr[c].x; N[c].y += fr[c].y; }
        if (a.a99.variant == A99_GPU && ix == 0) N[0].y = N[1].y = N[2].y = (T)0;
      }
    }
  }
}

// The stage mode is a template parameter, the grid is (plane, kz) and element indices are 32-bit: no 64-bit division per
// mode and no run-time switch, and the Y / A operands of the stage update are requested before the RHS arithmetic.
// (Round 1 shipped a grid-stride form with three 64-bit divisions per mode and a run-time stage switch; on B200 this form
// is 23 % faster -- 0.713 -> 0.550 ms at 256^3, 36.0 -> 27.6 ms per step at 1024^3, profiles/README.md -- and replaced it.
// The stage updates are written with explicit fused multiply-adds so the result does not depend on the compiler's
// contraction choices.)  Needs n_fields * field < 2^32 elements (true up to 1024^3 with 15 product fields).
template <typename T, int PHYS, int MODE, bool A99 = false, int VP = 0>
__global__ void __launch_bounds__(256, MHDF_SPEC_MINB) k_spectral(SpecArgs<T> a) {
  using C = Cx<T>;
  const SpecGeom<T>& g = a.g;
  constexpr int F = (PHYS == PHYS_MHD) ? 6 : 3;
  const unsigned plane = (unsigned)g.Kxp * (unsigned)g.Kyl;
  const unsigned e2 = blockIdx.x * 256u + threadIdx.x;
  if (e2 >= plane) return;
  const unsigned jc = e2 / (unsigned)g.Kxp, ix = e2 - jc * (unsigned)g.Kxp;
  if (ix >= (unsigned)g.Kx || g.ky0 + (int)jc >= g.by.count()) return;
  const int kc = blockIdx.y;
  const unsigned e = (unsigned)kc * plane + e2;
  const unsigned fld = (unsigned)g.field;
  // operands of the stage update first: their latency overlaps the RHS arithmetic
  C y[F], ac[F];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if constexpr (MODE == STEP_RK4_1 || MODE == STEP_RK4_2 || MODE == STEP_RK4_3) y[f] = a.Y[f * fld + e];
    if constexpr (MODE == STEP_RK4_2 || MODE == STEP_RK4_3 || MODE == STEP_RK4_4) ac[f] = a.A[f * fld + e];
    if constexpr (MODE == STEP_LSRK) ac[f] = a.first ? mk<C>(0, 0) : a.A[f * fld + e];
  }
  C N[F], sin[F];
  spec_rhs<T, PHYS, A99, VP>(a, e, (int)ix, (int)jc, kc, N, sin);
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const unsigned o = f * fld + e;
    if constexpr (MODE == STEP_CALCN) {
      a.Nout[o] = N[f];
    } else if constexpr (MODE == STEP_RK4_1) {   // Sin == Y
      a.A[o] = axpy(a.ca, N[f], y[f]);
      a.Sout[o] = axpy(a.cs, N[f], y[f]);
    } else if constexpr (MODE == STEP_RK4_2 || MODE == STEP_RK4_3) {
      a.A[o] = axpy(a.ca, N[f], ac[f]);
      a.Sout[o] = axpy(a.cs, N[f], y[f]);
    } else if constexpr (MODE == STEP_RK4_4) {
      a.Sout[o] = axpy(a.ca, N[f], ac[f]);
    } else {                                     // LSRK54: S2 = A_i S2 + dt N ; sol += B_i S2
      C s2 = mk<C>(a.dt * N[f].x, a.dt * N[f].y);
      if (!a.first) s2 = axpy(a.ca, ac[f], s2);
      a.A[o] = s2;
      a.Sout[o] = axpy(a.cs, s2, sin[f]);
    }
  }
}

// EMHD: derive the 18 inverse-transform inputs from B^ (compact):
//   A = i k x B (0..2), d_j B_i = i k_j B_i for j = y, z (3 + 2i + (j-1)), d_j A_i = i k_j A_i for j = y, z (9 + 2i + (j-1)),
//   B (15..17); the x derivatives are formed inside the fused x kernel from the rows of B_i / A_i
//   (reference: MHDSolver.jl:301-309 "way 2", :246, :257)
template <typename T>
__global__ void __launch_bounds__(256) k_emhd_derive(SpecGeom<T> g, const Cx<T>* __restrict__ B, Cx<T>* __restrict__ out) {
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    const T k[3] = {g.kx[ix], g.ky[jc], g.kz[kc]};
    C b[3], A[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = B[i * g.field + e];
    // A = i (k x B)
    const C c0 = mk<C>(k[1] * b[2].x - k[2] * b[1].x, k[1] * b[2].y - k[2] * b[1].y);
    const C c1 = mk<C>(k[2] * b[0].x - k[0] * b[2].x, k[2] * b[0].y - k[0] * b[2].y);
    const C c2 = mk<C>(k[0] * b[1].x - k[1] * b[0].x, k[0] * b[1].y - k[1] * b[0].y);
    A[0] = cmuli(c0); A[1] = cmuli(c1); A[2] = cmuli(c2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      out[i * g.field + e] = A[i];
      out[(15 + i) * g.field + e] = b[i];
#pragma unroll
      for (int j = 1; j < 3; ++j) {
        out[(3 + 2 * i + (j - 1)) * g.field + e] = cmuli(cscale(b[i], k[j]));
        out[(9 + 2 * i + (j - 1)) * g.field + e] = cmuli(cscale(A[i], k[j]));
      }
    }
  }
}

// DivVCorrection! / DivBCorrection! (Solver/VPSolver.jl:61-137): Phi = -i (k.f^) / k^2, f^_i -= i k_i Phi on three
// consecutive fields of a compact state, same operation order as the reference's broadcasts.
template <typename T>
__global__ void __launch_bounds__(256) k_divclean(SpecGeom<T> g, Cx<T>* __restrict__ S) {
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    const T k[3] = {g.kx[ix], g.ky[jc], g.kz[kc]};
    const T k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
    const T ik2 = (k2 > (T)0) ? (T)1 / k2 : (T)0;   // invKrsq[1,1,1] = 0
    C f[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) f[i] = S[i * g.field + e];
    const C s = mk<C>((k[0] * f[0].x + k[1] * f[1].x + k[2] * f[2].x) * ik2, (k[0] * f[0].y + k[1] * f[1].y + k[2] * f[2].y) * ik2);
#pragma unroll
    for (int i = 0; i < 3; ++i) S[i * g.field + e] = mk<C>(f[i].x - k[i] * s.x, f[i].y - k[i] * s.y);
  }
}

// ---- DivFreeSpectraMap (utils/IC.jl:130-179) on the device ---------------------------------------------------------------
// Random-phase power-law solenoidal field:  F^_i = A k^k0 e^{2 pi i theta} e2_i(k),  e2 = (kx kz, ky kz, -kp^2) / (kp k),
// kp^2 = kx^2 + ky^2, 0/0 -> 0, zero on the kr = 0 plane and for k < k_peak;  A = sqrt(3 P (Lx/dx)(Ly/dy)(Lz/dz) / sum(Fk/(k+1)^2)
// / dV) with the sum over the WHOLE (nkr, nl, nm) array (taken before dealias!, like the reference).  The reference draws
// theta = rand(T, nkr, nl, nm) from Julia's stream, which cannot be reproduced: theta is word 0 of the Philox4x32-10 block
// (key = seed, counter = index of the mode in the (nkr, nl, nm) array, tag word), the generator of the A99 driving.
// Arithmetic in T (k, Fk, the e2 basis as T broadcasts like IC.jl:139-159, evaluated left to right); e^{i theta 2 pi} in
// Float64 then rounded (IC.jl:163: `im .* rand(T, ...) * 2pi` promotes to Float64).
template <typename T>
struct DfsmArgs {
  int nkr, ny, nz;
  double dkx, dky, dkz;    // 2 pi / L per axis
  T k0, kpeak, amp;
  unsigned seed_lo, seed_hi;
};
enum : unsigned { DFSM_CALL_LO = 0x44465350u, DFSM_CALL_HI = 0x7FFFFFFFu };   // counter tag: never a forcing-call number
template <typename T>
__device__ __forceinline__ T dfsm_fk(T kx, T ky, T kz, int ix, T k0, T kpeak, T& k) {
  const T k2 = kx * kx + ky * ky + kz * kz;
  k = sqrt(k2);
  if (ix == 0 || !(k2 > (T)0) || k < kpeak) return (T)0;   // Fk[1,1,1] = 0; Fk[1,:,:] .= 0; Fk[k .< k_peak] .= 0
  return pow(k, k0);
}
// sum over every mode of the (nkr, ny, nz) array of Fk / (k + 1)^2  ->  out[0]  (Float64 accumulation)
template <typename T>
__global__ void __launch_bounds__(256) k_dfsm_norm(DfsmArgs<T> q, double* __restrict__ out) {
  const long long total = (long long)q.nkr * q.ny * q.nz;
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % q.nkr);
    const long long r = e / q.nkr;
    const int j = (int)(r % q.ny), kk = (int)(r / q.ny);
    const T kx = (T)(ix * q.dkx), ky = (T)((j < q.ny / 2 ? j : j - q.ny) * q.dky), kz = (T)((kk < q.nz / 2 ? kk : kk - q.nz) * q.dkz);
    T k;
    const T Fk = dfsm_fk<T>(kx, ky, kz, ix, q.k0, q.kpeak, k);
    const T k1 = k + (T)1;
    s += (double)(Fk * ((T)1 / (k1 * k1)));
  }
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_red_sum<double>(s);
  if (lane == 0) sh[wid] = s;
  __syncthreads();
  if (wid == 0) {
    double x = lane < nw ? sh[lane] : 0.0;
    warp_red_sum<double>(x);
    if (lane == 0) atomicAdd(out, x);
  }
}
// the three components on the retained modes of a compact state (fields S, S + field, S + 2 field)
template <typename T>
__global__ void __launch_bounds__(256) k_dfsm_fill(SpecGeom<T> g, DfsmArgs<T> q, Cx<T>* __restrict__ S) {
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  A99Args<T> rq;
  rq.seed_lo = q.seed_lo; rq.seed_hi = q.seed_hi; rq.call_lo = DFSM_CALL_LO; rq.call_hi = DFSM_CALL_HI;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    const int jg = g.ky0 + jc;
    if (jg >= g.by.count()) continue;
    const T kx = g.kx[ix], ky = g.ky[jc], kz = g.kz[kc];
    T k;
    const T Fk = dfsm_fk<T>(kx, ky, kz, ix, q.k0, q.kpeak, k) * q.amp;
    const T k2 = kx * kx + ky * ky + kz * kz;
    const T kinv = sqrt((k2 > (T)0) ? (T)1 / k2 : (T)0);
    const T kp = sqrt(kx * kx + ky * ky);
    T e2[3] = {kx * kz / kp * kinv, ky * kz / kp * kinv, -kp * kinv};
    if (e2[0] != e2[0]) e2[0] = (T)0;
    if (e2[1] != e2[1]) e2[1] = (T)0;
    const unsigned iy = (unsigned)(jg < g.by.lo ? jg : jg + (g.by.hi0 - g.by.lo));
    const unsigned iz = (unsigned)(kc < g.bz.lo ? kc : kc + (g.bz.hi0 - g.bz.lo));
    const unsigned long long mode = (unsigned)ix + (unsigned long long)q.nkr * (iy + (unsigned long long)g.by.n * iz);
    T r[4];
    a99_uniforms<T>(rq, mode, r);
    double sn, cs;
    sincospi(2.0 * (double)r[0], &sn, &cs);
    const T er = (T)cs, ei = (T)sn;
#pragma unroll
    for (int c = 0; c < 3; ++c) S[c * g.field + e] = mk<C>((Fk * er) * e2[c], (Fk * ei) * e2[c]);
  }
}

// ---- on-device analysis of the state (utils/MHDAnalysis.jl) ------------------------------------------------------------
// mode 0: ScaleDecomposition (MHDAnalysis.jl:24-82): f^_c <- f^_c where k1 <= |k| <= k2, else 0  (|k| = sqrt(kr^2 + l^2 + m^2) in T)
// mode 1: VectorPotential (MHDAnalysis.jl:129-174): a^ = i (k x b^) / k^2 (Coulomb gauge; the k = 0 mode gives 0)
// mode 2: power spectra |f^_c|^2 (real): their inverse transform is the autocorrelation CF(V) = real(ifft(|fft(V)|^2)) of each
//         component (utils/TurbStatTool.jl:67), the building block of the two-point structure functions SFC / SF_2 1D (:72, 90-120)
// in: three consecutive fields of a compact state; out: three compact fields (inverse-transformed by the caller).
template <typename T>
__global__ void __launch_bounds__(256) k_analysis(SpecGeom<T> g, const Cx<T>* __restrict__ S, int f0, Cx<T>* __restrict__ out, int mode, T k1, T k2) {
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    C f[3], o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { f[c] = S[(f0 + c) * g.field + e]; o[c] = mk<C>(0, 0); }
    if (ix < g.Kx && g.ky0 + jc < g.by.count()) {
      const T k[3] = {g.kx[ix], g.ky[jc], g.kz[kc]};
      const T kk2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
      if (mode == 0) {
        const T kr = sqrt(kk2);
        if (k2 >= kr && kr >= k1) { o[0] = f[0]; o[1] = f[1]; o[2] = f[2]; }
      } else if (mode == 2) {
        // kr = 0 plane: the power of what rfft of the REAL field holds there (the symmetrised mode); a mode whose mirror is dealiased
        // away is a real-space wave of half its amplitude on both sides: a quarter of its power each, i.e. 2 |f^sym|^2 = |f^|^2 / 2
        // here, which the Hermitian part taken by the inverse transform splits over the two
        bool unpaired = false;
        if (ix == 0) unpaired = g.by.row_of_wave(-g.by.wave(g.ky0 + jc)) < 0 || g.bz.row_of_wave(-g.bz.wave(kc)) < 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const C v = (ix == 0) ? load_sym<T>(S, f0 + c, g, ix, jc, kc) : f[c];
          const T p = v.x * v.x + v.y * v.y;
          o[c] = mk<C>(unpaired ? p + p : p, (T)0);
        }
      } else {
        const T ik2 = (kk2 > (T)0) ? (T)1 / kk2 : (T)0;
        const C c0 = mk<C>(k[1] * f[2].x - k[2] * f[1].x, k[1] * f[2].y - k[2] * f[1].y);
        const C c1 = mk<C>(k[2] * f[0].x - k[0] * f[2].x, k[2] * f[0].y - k[0] * f[2].y);
        const C c2 = mk<C>(k[0] * f[1].x - k[1] * f[0].x, k[0] * f[1].y - k[1] * f[0].y);
        o[0] = cscale(cmuli(c0), ik2); o[1] = cscale(cmuli(c1), ik2); o[2] = cscale(cmuli(c2), ik2);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c * g.field + e] = o[c];
  }
}

// full (nkr, ny, nz) spectral array <-> compact field.  dir = 0: full -> compact (drop dealiased modes),
// dir = 1: compact -> full (dealiased modes written as zero).
template <typename T>
__global__ void __launch_bounds__(256) k_pack(Cx<T>* __restrict__ full, Cx<T>* __restrict__ comp, int nkr, int ny, int nz,
                                              int Kx, int Kxp, Band by, Band bz, int dir, int ydirect) {
  // ydirect (slab runs): the host array is (nkr, Kyl, nz) -- its y rows ARE the local compact rows (ny == Kyl)
  using C = Cx<T>;
  const int Ky = ydirect ? ny : by.count();
  const long long total = (long long)nkr * ny * nz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % nkr);
    const long long r = e / nkr;
    const int j = (int)(r % ny), k = (int)(r / ny);
    const int jc = ydirect ? j : by.row(j), kc = bz.row(k);
    const bool kept = ix < Kx && jc >= 0 && kc >= 0;
    if (dir == 0) {
      if (kept) comp[((long long)kc * Ky + jc) * Kxp + ix] = full[e];
    } else {
      full[e] = kept ? comp[((long long)kc * Ky + jc) * Kxp + ix] : mk<C>(0, 0);
    }
  }
}

// Fresh diagnostics from the compact state via Parseval (kr = 0 plane symmetrised, weight 2 for kr > 0):
//   out[0] = sum |u|^2, out[1] = sum |b|^2, out[2] = sum u.(curl u), out[3] = sum a.b (Coulomb gauge),
//   out[4] = sum u.b      -- all as real-space sums over grid points (multiply by dV outside where wanted)
//   (reference: UserInterface.jl:29,65-86; MHDAnalysis.jl:94-117,165-168)
template <typename T>
__global__ void __launch_bounds__(256) k_diag(SpecGeom<T> g, const Cx<T>* __restrict__ S, int has_u, int has_b, int boff,
                                              double inv_n3, double* __restrict__ out) {
  // S = state base; velocity fields 0..2 if has_u, magnetic fields boff..boff+2 if has_b
  const Cx<T>* U = has_u ? S : nullptr;
  const Cx<T>* B = has_b ? S : nullptr;
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  double s[5] = {0, 0, 0, 0, 0};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    if (g.ky0 + jc >= g.by.count()) continue;
    const double k[3] = {(double)g.kx[ix], (double)g.ky[jc], (double)g.kz[kc]};
    const double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
    const double w = (ix == 0 ? 1.0 : 2.0) * inv_n3;
    double ur[3] = {0, 0, 0}, ui[3] = {0, 0, 0}, br[3] = {0, 0, 0}, bi[3] = {0, 0, 0};
    if (U != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { const C v = load_sym<T>(U, c, g, ix, jc, kc); ur[c] = v.x; ui[c] = v.y; }
      s[0] += w * (ur[0] * ur[0] + ui[0] * ui[0] + ur[1] * ur[1] + ui[1] * ui[1] + ur[2] * ur[2] + ui[2] * ui[2]);
      // omega = i k x u ; Re(u . conj(omega)) = sum_c Re(u_c conj(i c_c)) with c = k x u:  Re(u conj(i c)) = u_i c_r - u_r c_i
      const double cr[3] = {k[1] * ur[2] - k[2] * ur[1], k[2] * ur[0] - k[0] * ur[2], k[0] * ur[1] - k[1] * ur[0]};
      const double ci[3] = {k[1] * ui[2] - k[2] * ui[1], k[2] * ui[0] - k[0] * ui[2], k[0] * ui[1] - k[1] * ui[0]};
#pragma unroll
      for (int c = 0; c < 3; ++c) s[2] += w * (ui[c] * cr[c] - ur[c] * ci[c]);
    }
    if (B != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { const C v = load_sym<T>(B, boff + c, g, ix, jc, kc); br[c] = v.x; bi[c] = v.y; }
      s[1] += w * (br[0] * br[0] + bi[0] * bi[0] + br[1] * br[1] + bi[1] * bi[1] + br[2] * br[2] + bi[2] * bi[2]);
      if (k2 > 0) {
        // a = i (k x b) / k^2 ; Re(a . conj(b)) = sum_c Re(i c_c conj(b_c)) / k^2 = (c_r b_i - c_i b_r) / k^2
        const double cr[3] = {k[1] * br[2] - k[2] * br[1], k[2] * br[0] - k[0] * br[2], k[0] * br[1] - k[1] * br[0]};
        const double ci[3] = {k[1] * bi[2] - k[2] * bi[1], k[2] * bi[0] - k[0] * bi[2], k[0] * bi[1] - k[1] * bi[0]};
#pragma unroll
        for (int c = 0; c < 3; ++c) s[3] += w * (cr[c] * bi[c] - ci[c] * br[c]) / k2;
      }
      if (U != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s[4] += w * (ur[c] * br[c] + ui[c] * bi[c]);
      }
    }
  }
  {
    __shared__ double sh[32][5];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 5; ++i) warp_red_sum<double>(s[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) sh[wid][i] = s[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double x = lane < nw ? sh[lane][i] : 0.0;
        warp_red_sum<double>(x);
        if (lane == 0) atomicAdd(&out[i], x);
      }
    }
  }
}

// ---- HM89TimeStepper (timestepper/HM89.jl:23-199): the elementwise stages on three compact fields ------------------------
// HM_STAGE  one stage of LSRK3substeps! / RK3linearterm! (HM89.jl:141-199), in the order of the reference's broadcasts:
//             [lin: F -= eta k^2 S]   F *= dt   [G: F -= b G]   S += (F c) c2
//           eta, b are Float64 in the reference (params.eta; the literals 5/9, 153/128): those broadcasts run in Float64 and round on
//           the store; c = Float32/64 of the rational 1//3, 15//16, 8//15; c2 = dt in the first stage (HM89.jl:158,186), else 1
// HM_HALF   S = (B0 + B1) 0.5                                                   (HM89.jl:67)
// HM_FIXED  Bn = B0 + dt F1;  F = Bn - B1;  B1 = Bn                             (HM89.jl:71-82; Bn needs no array of its own,
//           and dealias!(Bn) is the compact layout itself)
enum { HM_STAGE = 0, HM_HALF = 1, HM_FIXED = 2 };
template <typename T>
struct Hm89Args {
  int op, lin;
  Cx<T>* F;          // STAGE: the stage array (F0 / F1); FIXED: out, B^n - B^1
  const Cx<T>* G;    // STAGE: the other stage array or null; FIXED: F1 = the Hall term of this iteration
  Cx<T>* S;          // STAGE, HALF: sol
  const Cx<T>* B0;
  Cx<T>* B1;
  double eta, b;
  T dt, c, c2;
};
template <typename T>
__global__ void __launch_bounds__(256) k_hm89(SpecGeom<T> g, Hm89Args<T> a) {
  using C = Cx<T>;
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    double ek2 = 0;
    if (a.op == HM_STAGE && a.lin) {
      const long long rowi = e / g.Kxp;
      const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
      const T kx = g.kx[ix], ky = g.ky[jc], kz = g.kz[kc];
      ek2 = a.eta * (double)(kx * kx + ky * ky + kz * kz);      // eta * k^2: Float64 * T
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const long long i = f * g.field + e;
      if (a.op == HM_STAGE) {
        C F = a.F[i];
        const C S = a.S[i];
        if (a.lin) F = mk<C>((T)((double)F.x - ek2 * (double)S.x), (T)((double)F.y - ek2 * (double)S.y));
        F = mk<C>(F.x * a.dt, F.y * a.dt);
        if (a.G != nullptr) {
          const C G = a.G[i];
          F = mk<C>((T)((double)F.x - a.b * (double)G.x), (T)((double)F.y - a.b * (double)G.y));
        }
        a.F[i] = F;
        a.S[i] = mk<C>(S.x + (F.x * a.c) * a.c2, S.y + (F.y * a.c) * a.c2);
      } else if (a.op == HM_HALF) {
        const C p = a.B0[i], q = a.B1[i];
        a.S[i] = mk<C>((p.x + q.x) * (T)0.5, (p.y + q.y) * (T)0.5);
      } else {
        const C p = a.B0[i], n = a.G[i], q = a.B1[i];
        const C bn = mk<C>(p.x + a.dt * n.x, p.y + a.dt * n.y);
        a.F[i] = mk<C>(bn.x - q.x, bn.y - q.y);
        a.B1[i] = bn;
      }
    }
  }
}

// square_mean of HM89substeps! (HM89.jl:45,79): max over the grid points of sqrt(x^2 + y^2 + z^2) of three real fields, in T;
// the bit pattern of the non-negative result goes to *out with an atomic max (see XRed::maxsq)
template <typename T>
__global__ void __launch_bounds__(256) k_norm3_max(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z, long long n,
                                                   unsigned long long* __restrict__ out) {
  T mx = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T a = x[i], b = y[i], c = z[i];
    T v = sqrt(a * a + b * b + c * c);
    if (!(v == v)) v = (T)INFINITY;                              // fmax drops NaNs: a NaN counts as +inf, the caller stops on it
    mx = max2(mx, v);
  }
  __shared__ T sh[8];
  warp_red_max(mx);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    T m = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = max2(m, sh[w]);
    atomic_max_bits(out, m);
  }
}

// Shell spectrum of one compact field: Pk[round(|k|)] += |f^|^2 over the HALF spectrum, no weights
// (reference: MHDAnalysis.jl:237-255 `spectralline`).  kr = 0 plane symmetrised (what rfft of the
// real field holds).  Shared-memory histogram per block, then one atomic per bin.
template <typename T>
__global__ void __launch_bounds__(256) k_spectrum(SpecGeom<T> g, const Cx<T>* __restrict__ S, int fi, double* __restrict__ Pk, int nbins) {
  using C = Cx<T>;
  MHDF_DYN_SMEM(double, hist);
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0.0;
  __syncthreads();
  const int Ky = g.Kyl, Kz = g.bz.count();
  const long long total = (long long)g.Kxp * Ky * Kz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(e % g.Kxp);
    if (ix >= g.Kx) continue;
    const long long rowi = e / g.Kxp;
    const int jc = (int)(rowi % Ky), kc = (int)(rowi / Ky);
    if (g.ky0 + jc >= g.by.count()) continue;
    const T kx = g.kx[ix], ky = g.ky[jc], kz = g.kz[kc];
    const T kk = sqrt(kx * kx + ky * ky + kz * kz);
    const int r = (int)rint((double)kk);
    const C v = load_sym<T>(S, fi, g, ix, jc, kc);
    if (r < nbins) atomicAdd(&hist[r], (double)v.x * v.x + (double)v.y * v.y);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (hist[i] != 0.0) atomicAdd(&Pk[i], hist[i]);
}

}  // namespace mhdf
