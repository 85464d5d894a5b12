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
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }

enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int bytes) { return bytes <= (int)sizeof(smem_raw) ? cudaSuccess : cudaErrorEmu; }

struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorEmu; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorEmu; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorEmu; }

// NCCL surface referenced by solver.cuh (bound with dlopen at run time; never reached with one rank)
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclUint32 = 3 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
