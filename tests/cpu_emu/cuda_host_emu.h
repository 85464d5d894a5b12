// cuda_host_emu.h -- host-side CUDA runtime / NCCL shim for the CPU test-suite ONLY (tests/test_emulated_library.py).
//
// With -DMHDF_CPU_EMU the whole library (api.cu + solver.cuh + the kernels) compiles as ordinary C++: device memory is the
// heap, streams and events are inert (everything runs synchronously), kernel launches go through emu::launch_call (one OS
// thread per CUDA thread, cuda_emu.h).  The resulting libmhdflows_b200_emu.so lives under the test tree, is never built by
// mhdflows_jl_b200.build and never shipped: it lets the host orchestration (buffer sizing, launch arguments, register
// rotation, API boundary) run on tiny grids in CI, nothing else.  Single rank only (no NCCL, no CUDA IPC).
#pragma once
#include "cuda_emu.h"

#include <cstdlib>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "CPU emulator: unsupported call"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; };
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 2; return cudaSuccess; }   // small grid-stride caps

template <typename U> inline cudaError_t cudaMalloc(U** p, size_t n) { *p = static_cast<U*>(std::malloc(n ? n : 1)); return *p ? cudaSuccess : cudaErrorEmu; }
template <typename U> inline cudaError_t cudaMallocHost(U** p, size_t n) { *p = static_cast<U*>(std::malloc(n ? n : 1)); return *p ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }

struct emuStream { int id; };
typedef emuStream* cudaStream_t;
struct emuEvent { int id; };
typedef emuEvent* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = new emuStream{0}; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }

enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int bytes) { return bytes <= (int)sizeof(smem_raw) ? cudaSuccess : cudaErrorEmu; }

// CUDA IPC inside one process (the multi-rank driver runs the ranks as threads): a handle is the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// NCCL surface referenced by solver.cuh.  The real library binds libnccl with dlopen; the emulated one binds the in-process
// stand-ins below (ranks = threads of tests/cpu_emu/test_library_ranks.cpp; every "stream" is synchronous, so a collective is a
// rendezvous of the rank threads).
struct ncclUniqueId { char internal[128]; };
typedef enum { ncclSuccess = 0, ncclEmuError = 1 } ncclResult_t;
typedef enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclUint32 = 3, ncclUint64 = 5 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
namespace emu_nccl {
struct SendOp { int peer; const void* p; size_t bytes; };
struct World {
  int n;
  std::unique_ptr<std::barrier<>> bar;
  std::vector<const void*> ptr;
  std::vector<std::vector<SendOp>> sends;
  explicit World(int n_) : n(n_), bar(std::make_unique<std::barrier<>>(n_)), ptr(n_), sends(n_) {}
};
struct Comm { World* w; int rank; };
inline std::mutex mu;
inline std::vector<std::pair<long long, World*>> worlds;
inline long long next_id = 1;
inline size_t size_of(ncclDataType_t t) { return (t == ncclFloat64 || t == ncclUint64) ? 8 : 4; }
struct GroupState { bool open = false; std::vector<SendOp> sends; struct R { int peer; void* p; size_t bytes; Comm* c; }; std::vector<R> recvs; Comm* c = nullptr; };
inline thread_local GroupState grp;
inline ncclResult_t GetUniqueId(ncclUniqueId* id) { std::lock_guard<std::mutex> g(mu); std::memset(id, 0, sizeof *id); const long long v = next_id++; std::memcpy(id->internal, &v, sizeof v); return ncclSuccess; }
inline ncclResult_t CommInitRank(Comm** c, int n, ncclUniqueId id, int rank) {
  long long v; std::memcpy(&v, id.internal, sizeof v);
  std::lock_guard<std::mutex> g(mu);
  World* w = nullptr;
  for (auto& e : worlds) if (e.first == v) w = e.second;
  if (!w) { w = new World(n); worlds.emplace_back(v, w); }
  *c = new Comm{w, rank};
  return ncclSuccess;
}
inline ncclResult_t CommDestroy(Comm* c) { delete c; return ncclSuccess; }
inline ncclResult_t AllReduce(const void* s, void* r, size_t count, ncclDataType_t t, ncclRedOp_t op, Comm* c, cudaStream_t) {
  World* w = c->w;
  w->ptr[c->rank] = s;
  w->bar->arrive_and_wait();
  std::vector<unsigned char> tmp(count * size_of(t));
  for (size_t i = 0; i < count; ++i) {
    if (t == ncclFloat64) { double a = 0; for (int q = 0; q < w->n; ++q) { const double x = ((const double*)w->ptr[q])[i]; a = op == ncclSum ? (q ? a + x : x) : (q ? std::max(a, x) : x); } ((double*)tmp.data())[i] = a; }
    else if (t == ncclFloat32) { float a = 0; for (int q = 0; q < w->n; ++q) { const float x = ((const float*)w->ptr[q])[i]; a = op == ncclSum ? (q ? a + x : x) : (q ? std::max(a, x) : x); } ((float*)tmp.data())[i] = a; }
    else if (t == ncclUint64) { unsigned long long a = 0; for (int q = 0; q < w->n; ++q) { const unsigned long long x = ((const unsigned long long*)w->ptr[q])[i]; a = op == ncclSum ? (q ? a + x : x) : (q ? std::max(a, x) : x); } ((unsigned long long*)tmp.data())[i] = a; }
    else { unsigned a = 0; for (int q = 0; q < w->n; ++q) { const unsigned x = ((const unsigned*)w->ptr[q])[i]; a = op == ncclSum ? (q ? a + x : x) : (q ? std::max(a, x) : x); } ((unsigned*)tmp.data())[i] = a; }
  }
  w->bar->arrive_and_wait();                       // everyone has read every send buffer (in-place calls)
  std::memcpy(r, tmp.data(), tmp.size());
  w->bar->arrive_and_wait();
  return ncclSuccess;
}
inline ncclResult_t AllGather(const void* s, void* r, size_t count, ncclDataType_t t, Comm* c, cudaStream_t) {
  World* w = c->w;
  w->ptr[c->rank] = s;
  w->bar->arrive_and_wait();
  const size_t b = count * size_of(t);
  for (int q = 0; q < w->n; ++q) std::memcpy((char*)r + (size_t)q * b, w->ptr[q], b);
  w->bar->arrive_and_wait();
  return ncclSuccess;
}
inline ncclResult_t GroupStart() { grp.open = true; grp.sends.clear(); grp.recvs.clear(); grp.c = nullptr; return ncclSuccess; }
inline ncclResult_t Send(const void* p, size_t count, ncclDataType_t t, int peer, Comm* c, cudaStream_t) { grp.c = c; grp.sends.push_back({peer, p, count * size_of(t)}); return grp.open ? ncclSuccess : ncclEmuError; }
inline ncclResult_t Recv(void* p, size_t count, ncclDataType_t t, int peer, Comm* c, cudaStream_t) { grp.c = c; grp.recvs.push_back({peer, p, count * size_of(t), c}); return grp.open ? ncclSuccess : ncclEmuError; }
inline ncclResult_t GroupEnd() {
  grp.open = false;
  if (!grp.c) return ncclSuccess;
  World* w = grp.c->w;
  w->sends[grp.c->rank] = grp.sends;
  w->bar->arrive_and_wait();
  for (auto& rv : grp.recvs)
    for (auto& sd : w->sends[rv.peer])
      if (sd.peer == grp.c->rank) { if (sd.bytes != rv.bytes) return ncclEmuError; std::memcpy(rv.p, sd.p, sd.bytes); break; }
  w->bar->arrive_and_wait();
  return ncclSuccess;
}
inline const char* GetErrorString(ncclResult_t) { return "emulated NCCL: mismatched group"; }
}  // namespace emu_nccl
typedef emu_nccl::Comm* ncclComm_t;
