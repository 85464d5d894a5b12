// fft_core.cuh -- register/shared-memory Stockham FFT building blocks for sm_100a.
//
// One FFT of length N is owned by Tn = N/E threads; thread t holds the E elements
// n = t + Tn*m (m = 0..E-1) in registers.  A transform is a sequence of radix-R steps
// (R = min(E, remaining), R in {2,4,8,16}); between steps the elements are exchanged
// through shared memory with the Stockham autosort index map
//     out(j, r') = (j / Ns) * Ns * R + (j % Ns) + r' * Ns ,   j = t + q*Tn .
// After the last step thread t again holds n = t + Tn*m in natural order, so a load /
// store pattern with TX (or Tn) contiguous elements per row segment is the same on both
// sides.  The index arithmetic is mirrored 1:1 in tools/fft_plan_sim.py.
//
// Replaces, for this path, the cuFFT/FFTW plans behind `mul!(yh, grid.rfftplan, y)` /
// `ldiv!(y, grid.rfftplan, yh)` (reference: src/Solver/MHDSolver.jl:74,93,152,167,333-338).
#pragma once
#ifdef MHDF_CPU_EMU
#include "cuda_emu.h"   // tests/cpu_emu: the same kernels compiled as plain C++ for the CPU test-suite
#else
#include <cuda_runtime.h>
#endif

namespace mhdf {

template <typename T> struct CxT;
template <> struct CxT<float>  { using type = float2; };
template <> struct CxT<double> { using type = double2; };
template <typename T> using Cx = typename CxT<T>::type;

template <typename C> struct RealOf;
template <> struct RealOf<float2>  { using type = float; };
template <> struct RealOf<double2> { using type = double; };

template <typename C> __device__ __forceinline__ C mk(typename RealOf<C>::type x, typename RealOf<C>::type y) { C c; c.x = x; c.y = y; return c; }

// ---- lane-wise pair arithmetic ---------------------------------------------------------------------------------
// A complex number (or two neighbouring real samples) is a pair; l* act on both lanes alike.  With -DMHDF_F32X2 the
// Float32 pairs map onto the sm_100 packed instructions add/sub/mul/fma.rn.f32x2 (SASS FADD2 / FMUL2 / FFMA2: one issue
// slot for both lanes, IEEE results identical to the scalar forms); ptxas folds the lane swaps, broadcasts and sign flips
// written below as plain `mk(...)` into operand modifiers (.LO_HI, .F32, -R, .NP).  Opt-in until measured on hardware
// (DESIGN.md section 7); the default build keeps the scalar forms.
template <typename C> __device__ __forceinline__ C ladd(C a, C b) { return mk<C>(a.x + b.x, a.y + b.y); }
template <typename C> __device__ __forceinline__ C lsub(C a, C b) { return mk<C>(a.x - b.x, a.y - b.y); }
template <typename C> __device__ __forceinline__ C lmul(C a, C b) { return mk<C>(a.x * b.x, a.y * b.y); }
// a * b + c and a * b - c * d per lane
template <typename C> __device__ __forceinline__ C lfma(C a, C b, C c) { return mk<C>(a.x * b.x + c.x, a.y * b.y + c.y); }
template <typename C> __device__ __forceinline__ C lmulsub(C a, C b, C c, C d) { return mk<C>(a.x * b.x - c.x * d.x, a.y * b.y - c.y * d.y); }
template <typename C> __device__ __forceinline__ C lneg(C a) { return mk<C>(-a.x, -a.y); }
template <typename C> __device__ __forceinline__ C lbc(typename RealOf<C>::type s) { return mk<C>(s, s); }

#ifdef MHDF_F32X2
#ifndef MHDF_CPU_EMU
__device__ __forceinline__ unsigned long long pk2_(float2 a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 up2_(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2 ladd(float2 a, float2 b) { unsigned long long r; asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2_(a)), "l"(pk2_(b))); return up2_(r); }
__device__ __forceinline__ float2 lsub(float2 a, float2 b) { unsigned long long r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2_(a)), "l"(pk2_(b))); return up2_(r); }
__device__ __forceinline__ float2 lmul(float2 a, float2 b) { unsigned long long r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2_(a)), "l"(pk2_(b))); return up2_(r); }
__device__ __forceinline__ float2 lfma(float2 a, float2 b, float2 c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2_(a)), "l"(pk2_(b)), "l"(pk2_(c))); return up2_(r); }
#else
// CPU emulation of the packed forms (same rounding: one fused multiply-add per lane)
inline float2 lfma(float2 a, float2 b, float2 c) { return mk<float2>(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
#endif
__device__ __forceinline__ float2 lmulsub(float2 a, float2 b, float2 c, float2 d) { return lfma(lneg(c), d, lmul(a, b)); }
#endif

// ---- float2p: a complex Float32 whose lane-wise arithmetic ALWAYS maps onto the packed sm_100 instructions -----------------
// Same layout as float2.  The strided axis passes instantiate the FFT building blocks with it (kernels.cuh: PassCx): measured
// on B200 the packed forms make those passes 3-5 % faster (profiles/r02_c1_ab_f32x2.log: fewer issue slots for a kernel that is
// co-limited by instruction issue and HBM), while the fused x kernel -- bound by shared-memory latency -- gains nothing from them
// at 512-point rows and stays on the scalar forms.  add / sub / mul are IEEE-identical per lane; a * b uses one product rounding
// and one fused multiply-add per lane.
struct __align__(8) float2p { float x, y; };
template <> struct RealOf<float2p> { using type = float; };
__device__ __forceinline__ float2p to_p(float2 a) { float2p r; r.x = a.x; r.y = a.y; return r; }
__device__ __forceinline__ float2 from_p(float2p a) { float2 r; r.x = a.x; r.y = a.y; return r; }
#ifndef MHDF_CPU_EMU
__device__ __forceinline__ unsigned long long pkp_(float2p a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2p upp_(unsigned long long v) { float2p r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2p ladd(float2p a, float2p b) { unsigned long long r; asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pkp_(a)), "l"(pkp_(b))); return upp_(r); }
__device__ __forceinline__ float2p lsub(float2p a, float2p b) { unsigned long long r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pkp_(a)), "l"(pkp_(b))); return upp_(r); }
__device__ __forceinline__ float2p lmul(float2p a, float2p b) { unsigned long long r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pkp_(a)), "l"(pkp_(b))); return upp_(r); }
__device__ __forceinline__ float2p lfma(float2p a, float2p b, float2p c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pkp_(a)), "l"(pkp_(b)), "l"(pkp_(c))); return upp_(r); }
#else
inline float2p lfma(float2p a, float2p b, float2p c) { return mk<float2p>(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
#endif

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { return ladd(a, b); }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { return lsub(a, b); }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { return mk<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) { return mk<C>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
template <typename C> __device__ __forceinline__ C cconj(C a) { return mk<C>(a.x, -a.y); }
template <typename C> __device__ __forceinline__ C cscale(C a, typename RealOf<C>::type s) { return lmul(a, lbc<C>(s)); }
// multiply by +i / -i
template <typename C> __device__ __forceinline__ C cmuli(C a)  { return mk<C>(-a.y, a.x); }
template <typename C> __device__ __forceinline__ C cmulmi(C a) { return mk<C>(a.y, -a.x); }
#ifdef MHDF_F32X2
// a * b = a.x * (b.x, b.y) + (-t.x, t.y),  t = a.y * (b.y, b.x);   a * conj(b) = a.y * (b.y, b.x) + (u.x, -u.y),  u = a.x * b:
// the one-lane sign flip sits on the addend, where FFMA2 has a per-lane negate modifier
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  const float2 t = lmul(lbc<float2>(a.y), mk<float2>(b.y, b.x));
  return lfma(lbc<float2>(a.x), b, mk<float2>(-t.x, t.y));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  const float2 u = lmul(lbc<float2>(a.x), b);
  return lfma(lbc<float2>(a.y), mk<float2>(b.y, b.x), mk<float2>(u.x, -u.y));
}
#endif

// a * b = a.x * (b.x, b.y) + (-t.x, t.y),  t = a.y * (b.y, b.x);   a * conj(b) = a.y * (b.y, b.x) + (u.x, -u.y),  u = a.x * b
__device__ __forceinline__ float2p cmul(float2p a, float2p b) {
  const float2p t = lmul(lbc<float2p>(a.y), mk<float2p>(b.y, b.x));
  return lfma(lbc<float2p>(a.x), b, mk<float2p>(-t.x, t.y));
}
__device__ __forceinline__ float2p cmulc(float2p a, float2p b) {
  const float2p u = lmul(lbc<float2p>(a.x), b);
  return lfma(lbc<float2p>(a.y), mk<float2p>(b.y, b.x), mk<float2p>(u.x, -u.y));
}

// DIR = -1: forward (exp(-i..)), DIR = +1: inverse (exp(+i..)), both unnormalised.

// cos/sin(2*pi*k/16), k = 0..7
__device__ __forceinline__ constexpr double c16(int k) {
  return k == 0 ? 1.0 : k == 1 ? 0.92387953251128673848 : k == 2 ? 0.70710678118654752440 : k == 3 ? 0.38268343236508977173
       : k == 4 ? 0.0 : k == 5 ? -0.38268343236508977173 : k == 6 ? -0.70710678118654752440 : -0.92387953251128673848;
}
__device__ __forceinline__ constexpr double s16(int k) {
  return k == 0 ? 0.0 : k == 1 ? 0.38268343236508977173 : k == 2 ? 0.70710678118654752440 : k == 3 ? 0.92387953251128673848
       : k == 4 ? 1.0 : k == 5 ? 0.92387953251128673848 : k == 6 ? 0.70710678118654752440 : 0.38268343236508977173;
}

// cos/sin(2*pi*k/32), k = 0..15 (radix-32 butterflies of the MHDF_PASS_E32 variant)
__device__ __forceinline__ constexpr double c32(int k) {
  return k == 0 ? 1.00000000000000000000 : k == 1 ? 0.98078528040323043058 : k == 2 ? 0.92387953251128673848 : k == 3 ? 0.83146961230254523567 : k == 4 ? 0.70710678118654757274 : k == 5 ? 0.55557023301960228867 : k == 6 ? 0.38268343236508983729 : k == 7 ? 0.19509032201612833135 : k == 8 ? 0.00000000000000006123 : k == 9 ? -0.19509032201612819257 : k == 10 ? -0.38268343236508972627 : k == 11 ? -0.55557023301960195560 : k == 12 ? -0.70710678118654746172 : k == 13 ? -0.83146961230254534669 : k == 14 ? -0.92387953251128673848 : -0.98078528040323043058;
}
__device__ __forceinline__ constexpr double s32(int k) {
  return k == 0 ? 0.00000000000000000000 : k == 1 ? 0.19509032201612824808 : k == 2 ? 0.38268343236508978178 : k == 3 ? 0.55557023301960217765 : k == 4 ? 0.70710678118654746172 : k == 5 ? 0.83146961230254523567 : k == 6 ? 0.92387953251128673848 : k == 7 ? 0.98078528040323043058 : k == 8 ? 1.00000000000000000000 : k == 9 ? 0.98078528040323043058 : k == 10 ? 0.92387953251128673848 : k == 11 ? 0.83146961230254545772 : k == 12 ? 0.70710678118654757274 : k == 13 ? 0.55557023301960217765 : k == 14 ? 0.38268343236508989280 : 0.19509032201612860891;
}

// In-register radix-R DFT, natural order in -> natural order out (decimation in time).
template <int R, int DIR, typename C> struct Bfly;

template <int DIR, typename C> struct Bfly<1, DIR, C> {
  static __device__ __forceinline__ void run(C (&)[1]) {}
};
template <int DIR, typename C> struct Bfly<2, DIR, C> {
  static __device__ __forceinline__ void run(C (&v)[2]) {
    C a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
  }
};
template <int DIR, typename C> struct Bfly<4, DIR, C> {
  static __device__ __forceinline__ void run(C (&v)[4]) {
    C a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    C a2 = cadd(v[1], v[3]), a3 = csub(v[1], v[3]);
    C r = (DIR < 0) ? cmulmi(a3) : cmuli(a3);   // a3 * exp(DIR*i*pi/2)
    v[0] = cadd(a0, a2);
    v[2] = csub(a0, a2);
    v[1] = cadd(a1, r);
    v[3] = csub(a1, r);
  }
};
template <int R, int DIR, typename C> struct Bfly {
  static_assert(R == 8 || R == 16 || R == 32, "radix must be 2, 4, 8, 16 or 32");
  using T = typename RealOf<C>::type;
  static __device__ __forceinline__ void run(C (&v)[R]) {
    C e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Bfly<R / 2, DIR, C>::run(e);
    Bfly<R / 2, DIR, C>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      C t;
      if (k == 0) t = o[0];
      else if (k == R / 4) t = (DIR < 0) ? cmulmi(o[k]) : cmuli(o[k]);
      else {
        const T wc = (T)(R == 32 ? c32(k) : c16(k * (16 / (R == 32 ? 16 : R))));
        const T ws = (T)(DIR * (R == 32 ? s32(k) : s16(k * (16 / (R == 32 ? 16 : R)))));
        t = cmul(o[k], mk<C>(wc, ws));
      }
      v[k] = cadd(e[k], t);
      v[k + R / 2] = csub(e[k], t);
    }
  }
};

constexpr __host__ __device__ int imin(int a, int b) { return a < b ? a : b; }

// Twiddle source.  tw[i] = (cos(2*pi*i/NTW), -sin(2*pi*i/NTW)), NTW = N * TWS (TWS = table stride, 1 or 2): read-only
// global table indexed by the twiddle exponent (a policy type so other sources can be plugged in; SLOT0 numbers the
// twiddles of a plan for such sources).
template <typename C> struct TwGlobal {
  const C* tw;
  __device__ __forceinline__ C get(int, unsigned gi) const { return __ldg(tw + gi); }
};
template <> struct TwGlobal<float2p> {
  const float2* tw;
  __device__ __forceinline__ float2p get(int, unsigned gi) const { return to_p(__ldg(tw + gi)); }
};

// One radix-R step on the register file of thread t (no shared memory).  SLOT0 = first twiddle slot of this step.
template <typename C, int N, int E, int R, int NS, int DIR, int TWS, int SLOT0, typename TW>
__device__ __forceinline__ void fft_step(C (&v)[E], int t, TW tw) {
  constexpr int Tn = N / E;
  constexpr int Q = E / R;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    C x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = v[q + r * Q];
    if (NS > 1) {
      const int j = t + q * Tn;
      constexpr unsigned STR = (N / (NS * R)) * TWS;
      const unsigned ks = (unsigned)(j & (NS - 1)) * STR;
#pragma unroll
      for (int r = 1; r < R; ++r) {
        C w = tw.get(SLOT0 + q * (R - 1) + (r - 1), (unsigned)r * ks);
        x[r] = (DIR < 0) ? cmul(x[r], w) : cmulc(x[r], w);
      }
    }
    Bfly<R, DIR, C>::run(x);
#pragma unroll
    for (int r = 0; r < R; ++r) v[q + r * Q] = x[r];
  }
}

// Scatter the outputs of the (R, NS) step into shared memory in Stockham order.
// IDX(n) maps the logical element index n of this FFT to a shared-memory slot.
template <typename C, int N, int E, int R, int NS, typename IDX>
__device__ __forceinline__ void fft_scatter(const C (&v)[E], int t, C* __restrict__ sm, IDX idx) {
  constexpr int Tn = N / E;
  constexpr int Q = E / R;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = t + q * Tn;
    const int base = (j / NS) * (NS * R) + (j & (NS - 1));
#pragma unroll
    for (int r = 0; r < R; ++r) sm[idx(base + r * NS)] = v[q + r * Q];
  }
}
template <typename C, int N, int E, typename IDX>
__device__ __forceinline__ void fft_gather(C (&v)[E], int t, const C* __restrict__ sm, IDX idx) {
  constexpr int Tn = N / E;
#pragma unroll
  for (int m = 0; m < E; ++m) v[m] = sm[idx(t + Tn * m)];
}

// SYNC policy objects: block-wide or warp-wide barrier.
struct SyncBlock { static __device__ __forceinline__ void sync() { __syncthreads(); } };
struct SyncWarp  { static __device__ __forceinline__ void sync() { __syncwarp(); } };

// Full transform.  Shared memory is used as two alternating exchange buffers (sm0, sm1) so each
// exchange costs one barrier: step s writes buffer (s & 1); a thread can only reach the next write
// of the same buffer after passing the barrier of the exchange in between, by which time every
// thread has finished reading it.  The caller must ensure sm0 is free on entry (one barrier since
// its last read).  Returns with data in registers in natural order n = t + Tn*m.
// idx maps element -> slot for the first exchange, idx2 for the later ones (each exchange writes and reads one buffer
// with one map, so the maps may differ: the scatter patterns of the first and the later steps need different padding
// to be bank-conflict free).
template <typename C, int N, int E, int DIR, int TWS, int NS, int PAR, typename SYNC, typename IDX, typename IDX2, typename TW, int SLOT0 = 0>
__device__ __forceinline__ void fft_run(C (&v)[E], int t, C* __restrict__ sm0, C* __restrict__ sm1, TW tw, IDX idx, IDX2 idx2) {
  constexpr int R = imin(E, N / NS);
  fft_step<C, N, E, R, NS, DIR, TWS, SLOT0>(v, t, tw);
  if constexpr (NS * R < N) {
    C* sm = PAR ? sm1 : sm0;
    if constexpr (NS == 1) {
      fft_scatter<C, N, E, R, NS>(v, t, sm, idx);
      SYNC::sync();
      fft_gather<C, N, E>(v, t, sm, idx);
    } else {
      fft_scatter<C, N, E, R, NS>(v, t, sm, idx2);
      SYNC::sync();
      fft_gather<C, N, E>(v, t, sm, idx2);
    }
    constexpr int NEXT = SLOT0 + (NS > 1 ? (E / R) * (R - 1) : 0);
    fft_run<C, N, E, DIR, TWS, NS * R, PAR ^ 1, SYNC, IDX, IDX2, TW, NEXT>(v, t, sm0, sm1, tw, idx, idx2);
  }
}

// Single-buffer variant (strided passes, where shared memory limits occupancy): two barriers per
// exchange.  The buffer must be free on entry; it is free again on return.
template <typename C, int N, int E, int DIR, int TWS, int NS, typename IDX, typename TW, int SLOT0 = 0>
__device__ __forceinline__ void fft_run_sb(C (&v)[E], int t, C* __restrict__ sm, TW tw, IDX idx) {
  constexpr int R = imin(E, N / NS);
  fft_step<C, N, E, R, NS, DIR, TWS, SLOT0>(v, t, tw);
  if constexpr (NS * R < N) {
    if constexpr (NS > 1) __syncthreads();   // previous gather finished
    fft_scatter<C, N, E, R, NS>(v, t, sm, idx);
    __syncthreads();
    fft_gather<C, N, E>(v, t, sm, idx);
    constexpr int NEXT = SLOT0 + (NS > 1 ? (E / R) * (R - 1) : 0);
    fft_run_sb<C, N, E, DIR, TWS, NS * R, IDX, TW, NEXT>(v, t, sm, tw, idx);
  }
}

// number of exchanges (shared-memory round trips) of a plan
constexpr __host__ __device__ int fft_num_steps(int N, int E) {
  int s = 0, ns = 1;
  while (ns < N) { ns *= imin(E, N / ns); ++s; }
  return s;
}

}  // namespace mhdf
