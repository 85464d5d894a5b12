// cuda_emu.h -- minimal CUDA-on-CPU shim used ONLY by the CPU test-suite (tests/test_kernel_emulation.py).
//
// The kernels of mhdflows_jl_b200/csrc/*.cuh are compiled as ordinary C++ (g++ -DMHDF_CPU_EMU): every CUDA thread of a
// block becomes an OS thread, __syncthreads / __syncwarp are real barriers, warp shuffles exchange through a per-warp
// scratch line, blocks run one after another.  Good for tiny grids (16^3 .. 64-point axes): it checks the index
// arithmetic, barrier placement and shuffle patterns of the very same source the GPU runs -- not performance.
#pragma once
#include <vector_types.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __restrict__
#undef __launch_bounds__
#undef __align__
#undef __shared__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {
struct Warp {
  std::unique_ptr<std::barrier<>> bar;
  double scratch[32][2];
};
struct Block {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<Warp> warps;
};
inline thread_local Block* cur_block = nullptr;
inline thread_local int lane_id = 0, warp_id = 0;
inline std::mutex atomic_mu;
inline std::mutex launch_mu;   // one kernel launch at a time, process-wide (not per launch_call instantiation)
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

// dynamic shared memory of the running block (blocks run one at a time)
alignas(16) inline unsigned char smem_raw[256 * 1024];
alignas(16) inline double hist[8192];
#define __shared__ static
// `extern __shared__ T name[];` inside a kernel -> refers to the global arrays above
#define MHDF_EMU_EXTERN_SHARED 1

inline void __syncthreads() { emu::cur_block->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur_block->warps[emu::warp_id].bar->arrive_and_wait(); }

template <typename V> inline V __shfl_sync(unsigned, V v, int src, int = 32) {
  emu::Warp& w = emu::cur_block->warps[emu::warp_id];
  std::memcpy(&w.scratch[emu::lane_id][0], &v, sizeof(V));
  w.bar->arrive_and_wait();
  V r;
  std::memcpy(&r, &w.scratch[src & 31][0], sizeof(V));
  w.bar->arrive_and_wait();
  return r;
}
template <typename V> inline V __shfl_xor_sync(unsigned m, V v, int x, int = 32) { return __shfl_sync(m, v, emu::lane_id ^ x); }

template <typename V> inline V __ldg(const V* p) { return *p; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline double atomicAdd(double* p, double v) { std::lock_guard<std::mutex> g(emu::atomic_mu); double o = *p; *p = o + v; return o; }
inline unsigned atomicMax(unsigned* p, unsigned v) { std::lock_guard<std::mutex> g(emu::atomic_mu); unsigned o = *p; *p = std::max(o, v); return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { std::lock_guard<std::mutex> g(emu::atomic_mu); unsigned long long o = *p; *p = std::max(o, v); return o; }
inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
using std::fmaxf;
using std::fmax;
using std::rint;
inline float __fmul_rn(float a, float b) { return a * b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }

namespace emu {
// run body() once per CUDA thread of a (gx, gy, gz) grid of 1-D blocks of `nthreads` threads, blocks one after another
template <typename F>
void launch_call(dim3 grid, int nthreads, F&& body) {
  // blocks run one at a time and share smem_raw / the function-local __shared__ arrays: launches of concurrent host threads
  // (the rank threads of the multi-rank driver) take turns
  std::lock_guard<std::mutex> launch_guard(launch_mu);
  const int nw = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk;
        blk.bar = std::make_unique<std::barrier<>>(nthreads);
        blk.warps.resize(nw);
        for (int w = 0; w < nw; ++w) blk.warps[w].bar = std::make_unique<std::barrier<>>(std::min(32, nthreads - 32 * w));
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
          th.emplace_back([&, t] {
            cur_block = &blk;
            lane_id = t & 31;
            warp_id = t >> 5;
            threadIdx = uint3{(unsigned)t, 0, 0};
            blockIdx = uint3{bx, by, bz};
            blockDim = dim3(nthreads, 1, 1);
            gridDim = grid;
            body();
            // a thread that has left the kernel no longer takes part in barriers (CUDA semantics)
            blk.bar->arrive_and_drop();
            blk.warps[warp_id].bar->arrive_and_drop();
          });
        for (auto& x : th) x.join();
      }
}
// run kernel(args) on a grid
template <typename K, typename A>
void launch(K kernel, dim3 grid, int nthreads, const A& args) {
  launch_call(grid, nthreads, [&] { kernel(args); });
}
}  // namespace emu
