// TEST INFRASTRUCTURE ONLY. Lets g++ compile kart_b200/csrc/kb_api.cu for the host (-DKB_EMUL) so that the C ABI, the
// host-side logic and the per-thread bodies of the kernels can be exercised in the CPU-only CI container next to the oracle.
// The result (tests/emul/libkartb200_emul.so) is never shipped, never loaded by kart_b200/ and is not a fallback of the
// product: kart_b200/libkartb200.so contains only the CUDA path.
#ifndef KB_CUDA_SHIM_H
#define KB_CUDA_SHIM_H
#include <chrono>
#include <cstdlib>
#include <cstring>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
struct kb_dim3 { unsigned x, y, z; };
static thread_local kb_dim3 blockIdx, blockDim, threadIdx, gridDim;
typedef int cudaError_t;
typedef void* cudaStream_t;
struct kb_emul_event { std::chrono::steady_clock::time_point t; };
typedef kb_emul_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("KB_EMUL_DEVICES"); *n = e && atoi(e) > 0 ? atoi(e) : 1; return cudaSuccess; }   // KB_EMUL_DEVICES=N: the host's multi-device worker pool over N emulated devices
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new kb_emul_event(); return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = new kb_emul_event(); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class F> static inline void kb_emul_launch(unsigned grid, unsigned block, F body)
{
	gridDim.x = grid; blockDim.x = block;
	for (unsigned b = 0; b < grid; b++) for (unsigned t = 0; t < block; t++) { blockIdx.x = b; threadIdx.x = t; body(); }
}
#endif
