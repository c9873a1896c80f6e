// Random 32-byte sector reads from a buffer much larger than L2: the ceiling the seeding kernel's access pattern has on this GPU.
//   independent : every thread issues its loads back to back (addresses do not depend on loaded data) -> DRAM random-sector peak
//   dependent   : one chain per thread, the next address comes out of the sector just loaded (what one FM-index extension
//                 step does: Occ block -> new interval -> next Occ block) at the residency k_fm_seed runs with
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/randsector scripts/gpu_randsector.cu
// Run:   scripts/bin/randsector [buffer MiB = 3072]      (prints one JSON line)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t ld_sector(const uint32_t* p)
{
	uint32_t r[8];
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
	return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}
__global__ void k_fill(uint32_t* buf, size_t words) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) buf[i] = mix((uint32_t)i * 2654435761u + 12345u); }
__global__ void k_indep(const uint32_t* buf, uint32_t nsec, int iters, uint32_t* out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
#pragma unroll 8
	for (int i = 0; i < iters; i++) acc ^= ld_sector(buf + (size_t)(mix(t * 0x9E3779B9u + (uint32_t)i) % nsec) * 8);
	if (acc == 0x12345678u) out[0] = acc;
}
__global__ void k_dep(const uint32_t* buf, uint32_t nsec, int iters, uint32_t* out)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, x = mix(t + 1u);
	for (int i = 0; i < iters; i++) x = mix(x ^ ld_sector(buf + (size_t)(x % nsec) * 8));
	if (x == 0x12345678u) out[0] = x;
}

template <class K> static double run(K kern, const uint32_t* buf, uint32_t nsec, int blocks, int threads, int iters, uint32_t* out)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	kern<<<blocks, threads>>>(buf, nsec, iters, out);
	cudaEventRecord(a);
	kern<<<blocks, threads>>>(buf, nsec, iters, out);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	return (double)blocks * threads * iters * 32.0 / (ms * 1e-3) / 1e9;
}

int main(int argc, char** argv)
{
	size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 3072;
	size_t bytes = mib << 20, words = bytes / 4; uint32_t nsec = (uint32_t)(bytes / 32);
	uint32_t *buf, *out; if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&out, 64) != cudaSuccess) { fprintf(stderr, "no device memory\n"); return 1; }
	cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); int sm = pr.multiProcessorCount;
	k_fill<<<sm * 8, 256>>>(buf, words); cudaDeviceSynchronize();
	double indep = run(k_indep, buf, nsec, sm * 16, 128, 256, out);
	double d10 = run(k_dep, buf, nsec, sm * 10, 128, 200, out);     // 1280 chains per SM: k_fm_seed<10>
	double d12 = run(k_dep, buf, nsec, sm * 12, 128, 200, out);
	double d16 = run(k_dep, buf, nsec, sm * 16, 128, 200, out);     // 2048 chains per SM: the residency limit
	printf("{\"buffer_mib\": %zu, \"sms\": %d, \"independent_gbs\": %.1f, \"dependent_1280_per_sm_gbs\": %.1f, \"dependent_1536_per_sm_gbs\": %.1f, \"dependent_2048_per_sm_gbs\": %.1f}\n", mib, sm, indep, d10, d12, d16);
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
