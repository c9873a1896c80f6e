// GPU construction of the BWA-format FM-index of Kart's 2G text (forward strand + reverse complement) from the packed forward
// strand: the contents of the .bwt file (BWT with the Occ counts interleaved every 128 symbols) and of the .sa file (every 32nd
// suffix-array entry). Replaces, for `kart index -gpu`, what the reference's builder does in bwt_pac2bwt / BWTIncConstructFromPacked
// (src/BWT_Index/bwt_gen.c:1601), bwt_bwtupdate_core (bwtindex.c:53-75) and bwt_cal_sa (bwt.c:101-123) -- hours for a human-sized
// genome on one core -- and what this repo's host builder (host/index_build.cpp, minutes on 16 cores) does with a comparison sort.
// The files are canonical functions of the text, so the output is byte-identical to both (tests/test_gpu_index_build.py).
//
// Not a port: the reference inserts the text into the BWT block by block (BWT-SW). Here the suffix array itself is built in HBM --
// a B200 holds the 6.2 G suffix positions of a 3.1 Gbp genome (50 GB) next to the text -- by MSD radix sorting on 29-base keys:
//   * suffixes are cut into super-buckets by their first 6 bases (histogram, then ranges of at most ~2^29 suffixes);
//   * a super-bucket's suffixes are selected, keyed with their first 29 bases (58 bits) plus 6 bits "how many of those bases
//     exist" (a suffix that ends inside the key sorts in front of the ones that go on with A's: the end of the text is the
//     smallest symbol, as in the reference's BWT), and radix-sorted;
//   * runs of equal keys are refined with the next 29 bases, and so on, only over the suffixes that are still tied: per round
//     one sort by the new key and one stable sort by the run id, after which the k-th element belongs into the k-th tied slot;
//     runs die out geometrically on diverged repeats (the synthetic genome's repeat families need ~20 rounds).
// The device-wide sorts, scans and selections are CUB's (header-only, compiled into this library); the kernels around them are ours.
// Then: BWT symbol of row r = text[SA[r] - 1], the 128-symbol blocks with running counts, and the samples.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/kart_b200.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

namespace {

std::string g_err;
int fail(const char* what, cudaError_t e = cudaSuccess)
{
	g_err = what; if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
	return e == cudaErrorMemoryAllocation ? KB_ENOMEM : KB_ECUDA;
}
#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, e_); } while (0)

template <class T> struct Buf
{
	T* p = nullptr; size_t n = 0;
	cudaError_t get(size_t want) { if (want <= n) return cudaSuccess; drop(); cudaError_t e = cudaMalloc((void**)&p, (want ? want : 1) * sizeof(T)); if (e == cudaSuccess) n = want; return e; }
	void drop() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~Buf() { drop(); }
};

// ---- the 2G text: 32 bases per big-endian word (base i of a word at bits 62-2i), two zero words behind the end ----
__global__ void k_text_words(const u8* pac, u64 L, u64 N, u64 words, u64* T)
{
	const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= words) return;
	u64 v = 0;
	for (int i = 0; i < 32; i++)
	{
		const u64 p = 32 * w + i; if (p >= N) break;
		u32 c;
		if (p < L) c = (pac[p >> 2] >> ((~p & 3) << 1)) & 3u;
		else { const u64 f = N - 1 - p; c = 3u - ((pac[f >> 2] >> ((~f & 3) << 1)) & 3u); }   // bntseq.c:190-191
		v |= (u64)c << (62 - 2 * i);
	}
	T[w] = v;
}
__device__ __forceinline__ u64 text_win(const u64* T, u64 p)   // 32 bases from p (zeros behind the end of the text)
{
	const u64 i = p >> 5; const int s = (int)(p & 31) * 2;
	const u64 a = T[i];
	return s ? (a << s) | (T[i + 1] >> (64 - s)) : a;
}
// sort key of suffix p at depth d: 29 bases from p + d (58 bits), then how many of them exist (0..29)
__device__ __forceinline__ u64 key29(const u64* T, u64 N, u64 p, u64 d)
{
	const u64 q = p + d;
	if (q >= N) return 0;
	const u64 left = N - q; const u64 valid = left < 29 ? left : 29;
	return (text_win(T, q) & ~0x3Full) | valid;
}
__device__ __forceinline__ u32 base_at(const u64* T, u64 p) { return (u32)(T[p >> 5] >> (62 - 2 * (p & 31))) & 3u; }

// ---- histogram of the first 6 bases ----
#define KB_IB_BINS 4096
#define KB_IB_SHIFT 52   // the first 6 bases of a 32-base window
__global__ void k_hist7(const u64* T, u64 N, unsigned long long* hist)
{
	__shared__ u32 h[KB_IB_BINS];
	for (int i = threadIdx.x; i < KB_IB_BINS; i += blockDim.x) h[i] = 0;
	__syncthreads();
	const u64 per = 1u << 16, lo = (u64)blockIdx.x * per, hi = lo + per < N ? lo + per : N;
	for (u64 p = lo + threadIdx.x; p < hi; p += blockDim.x) atomicAdd(&h[(u32)(text_win(T, p) >> KB_IB_SHIFT)], 1u);
	__syncthreads();
	for (int i = threadIdx.x; i < KB_IB_BINS; i += blockDim.x) if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}
struct InRange   // selection predicate: suffix p starts with a 6-mer in [lo, hi)
{
	const u64* T; u32 lo, hi;
	__device__ __forceinline__ bool operator()(const u64& p) const { const u32 b = (u32)(text_win(T, p) >> KB_IB_SHIFT); return b >= lo && b < hi; }
};
__global__ void k_keys(const u64* T, u64 N, const u64* pos, u64 n, u64 d, u64* key)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) key[i] = key29(T, N, pos[i], d);
}
__global__ void k_keys_idx(const u64* T, u64 N, const u64* pos, u32 n, u64 d, u64* key, u32* idx)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { key[i] = key29(T, N, pos[i], d); idx[i] = i; }
}
// after the first sort of a super-bucket: positions into the suffix array, run heads, and which suffixes are still tied
__global__ void k_first_round(const u64* key, const u64* pos, u64 n, u64* sa_out, u32* head_slot, u8* active)
{
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	sa_out[i] = pos[i];
	const bool head = i == 0 || key[i] != key[i - 1];
	const bool next_head = i + 1 == n || key[i + 1] != key[i];
	head_slot[i] = head ? (u32)i : 0u;
	active[i] = (head && next_head) ? 0 : 1;
}
__global__ void k_gather_u32(const u32* src, const u32* idx, u32 n, u32* dst) { const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }
// after a refinement round: k-th element of the (run, key) order -> k-th tied slot; new run heads; who is still tied
__global__ void k_next_round(const u64* key_unsorted, const u64* pos_in, const u32* slot, const u32* gid_sorted, const u32* order, u32 n,
                             u64* sa_bucket, u64* pos_out, u32* head_slot, u8* active)
{
	const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const u32 me = order[k]; const u64 kk = key_unsorted[me]; const u32 g = gid_sorted[k];
	const u64 p = pos_in[me];
	sa_bucket[slot[k]] = p; pos_out[k] = p;
	bool head = k == 0, next_head = k + 1 == n;
	if (!head) head = gid_sorted[k - 1] != g || key_unsorted[order[k - 1]] != kk;
	if (!next_head) next_head = gid_sorted[k + 1] != g || key_unsorted[order[k + 1]] != kk;
	head_slot[k] = head ? slot[k] : 0u;
	active[k] = (head && next_head) ? 0 : 1;
}
struct MaxOp { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };

// ---- BWT, Occ interleave, samples ----
__global__ void k_find_primary(const u64* sa, u64 N, unsigned long long* primary) { const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (r < N && sa[r] == 0) *primary = r + 1; }
// one thread per 16 symbols: the symbol word at its place in the .bwt layout, and the word's base counts
__global__ void k_bwt_words(const u64* T, const u64* sa, u64 N, u64 primary, u32* out, u32* cnt4)
{
	const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x, nw = (N + 15) / 16;
	if (w >= nw) return;
	u32 v = 0, c[4] = {0, 0, 0, 0};
	for (int j = 0; j < 16; j++)
	{
		const u64 i = 16 * w + j; if (i >= N) break;
		u32 s;
		if (i == 0) s = base_at(T, N - 1);                                  // row 0: the empty suffix, preceded by the last base
		else { const u64 row = i < primary ? i : i + 1; s = base_at(T, sa[row - 1] - 1); }   // rows behind `primary` move up by one
		v |= s << ((15 - j) << 1); c[s]++;
	}
	const u64 blk = w >> 3;
	out[blk * 16 + 8 + (w & 7)] = v;
	cnt4[w] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
}
// per 128-symbol block: its base counts (for the scan)
__global__ void k_block_counts(const u32* cnt4, u64 nw, u64 nblk, u64* cA, u64* cC, u64* cG, u64* cT)
{
	const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblk) return;
	u64 a = 0, c = 0, g = 0, t = 0;
	for (int k = 0; k < 8; k++) { const u64 w = 8 * b + k; if (w >= nw) break; const u32 v = cnt4[w]; a += v & 255; c += (v >> 8) & 255; g += (v >> 16) & 255; t += v >> 24; }
	cA[b] = a; cC[b] = c; cG[b] = g; cT[b] = t;
}
// the running counts in front of every block (after an exclusive scan); entry nblk = the totals behind the last symbol word
__global__ void k_put_counts(const u64* cA, const u64* cC, const u64* cG, const u64* cT, u64 nblk, u64 tail_at, u32* out)
{
	const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nblk) return;
	// 32-bit halves: the totals behind a partial last block start at an odd word when that block holds an odd number of symbol words
	u32* o = out + (b < nblk ? b * 16 : tail_at);
	const u64 v[4] = {cA[b], cC[b], cG[b], cT[b]};
	for (int c = 0; c < 4; c++) { o[2 * c] = (u32)v[c]; o[2 * c + 1] = (u32)(v[c] >> 32); }
}
__global__ void k_samples(const u64* sa, u64 n_sa, u64* smp) { const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (j >= 1 && j < n_sa) smp[j - 1] = sa[j * 32 - 1]; }

inline unsigned grid_for(u64 n, unsigned block = 256) { return (unsigned)((n + block - 1) / block); }

}   // namespace

extern "C" {

const char* kb_index_build_error(void) { return g_err.c_str(); }

void kb_index_free(kb_built_index_t* b)
{
	if (!b) return;
	if (b->bwt) cudaFreeHost(b->bwt);
	if (b->sa) cudaFreeHost(b->sa);
	b->bwt = nullptr; b->sa = nullptr;
}

int kb_index_build(int device, const uint8_t* pac, int64_t l_pac, kb_built_index_t* out)
{
	if (!pac || l_pac <= 0 || !out) { g_err = "kb_index_build: bad arguments"; return KB_EINVAL; }
	memset(out, 0, sizeof(*out));
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_err = "no CUDA device (this library has no CPU path)"; return KB_ENODEV; }
	if (device < 0 || device >= ndev) { g_err = "kb_index_build: no such device"; return KB_EINVAL; }
	CKI(cudaSetDevice(device));
	const bool verbose = getenv("KB_INDEX_TRACE") != nullptr;
	cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, 0);
	auto lap = [&](const char* what) { if (!verbose) return; cudaEventRecord(ev1, 0); cudaEventSynchronize(ev1); float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); fprintf(stderr, "[kb index] %-28s %.1f ms since start\n", what, ms); };
	const u64 L = (u64)l_pac, N = 2 * L, words = N / 32 + 3;
	if (N + 64 >= (1ull << 40)) { g_err = "kb_index_build: text too long"; return KB_EINVAL; }

	// ---- text ----
	Buf<u8> d_pac; Buf<u64> T;
	const size_t pac_bytes = (size_t)(L / 4 + 1);
	CKI(d_pac.get(pac_bytes)); CKI(T.get(words));
	CKI(cudaMemcpy(d_pac.p, pac, pac_bytes, cudaMemcpyHostToDevice));
	CKI(cudaMemset(T.p, 0, words * 8));
	k_text_words<<<grid_for(words), 256>>>(d_pac.p, L, N, words - 2, T.p);
	CKI(cudaGetLastError());
	d_pac.drop();

	// ---- super-buckets ----
	Buf<unsigned long long> d_hist; CKI(d_hist.get(KB_IB_BINS)); CKI(cudaMemset(d_hist.p, 0, KB_IB_BINS * 8));
	k_hist7<<<(unsigned)((N + 65535) >> 16), 256>>>(T.p, N, d_hist.p);
	CKI(cudaGetLastError());
	std::vector<unsigned long long> hist(KB_IB_BINS);
	CKI(cudaMemcpy(hist.data(), d_hist.p, KB_IB_BINS * 8, cudaMemcpyDeviceToHost));
	lap("text + histogram");
	u64 target = 1ull << 29;
	{ const char* e = getenv("KB_INDEX_BUCKET"); if (e && atoll(e) >= 1024) target = (u64)atoll(e); }
	struct Range { u32 lo, hi; u64 count, base; };
	std::vector<Range> ranges; u64 cmax = 0;
	{
		u64 run = 0, base = 0; u32 lo = 0;
		for (u32 b = 0; b < KB_IB_BINS; b++)
		{
			if (run > 0 && run + hist[b] > target) { ranges.push_back({lo, b, run, base}); base += run; run = 0; lo = b; }
			run += hist[b];
		}
		ranges.push_back({lo, (u32)KB_IB_BINS, run, base});
		for (const Range& r : ranges) cmax = std::max(cmax, r.count);
		if (base + run != N) { g_err = "kb_index_build: histogram does not add up"; return KB_ECUDA; }
	}
	if (cmax >= 0xFFFFFFF0ull) { g_err = "kb_index_build: more than 2^32 suffixes share their first 6 bases; use the host builder for this genome"; return KB_ECAPACITY; }

	// ---- suffix array ----
	Buf<u64> SA; CKI(SA.get(N));
	Buf<u64> keyA, keyB, posA, posB; Buf<u32> slotA, slotB, gidA, gidB, idxA, idxB, headv; Buf<u8> active, temp; Buf<u64> d_count;
	CKI(keyA.get(cmax)); CKI(keyB.get(cmax)); CKI(posA.get(cmax)); CKI(posB.get(cmax));
	CKI(headv.get(cmax)); CKI(active.get(cmax)); CKI(d_count.get(2));
	size_t temp_bytes = 0;
	{
		// one temporary buffer for every CUB call below (queried at the largest size)
		size_t a = 0, b = 0, c = 0, d = 0, e = 0;
		thrust::counting_iterator<u64> it(0); InRange pr{T.p, 0, 1};
		cub::DeviceSelect::If(nullptr, a, it, posA.p, d_count.p, (::cuda::std::int64_t)N, pr);
		cub::DoubleBuffer<u64> dk(keyA.p, keyB.p), dv(posA.p, posB.p);
		cub::DeviceRadixSort::SortPairs(nullptr, b, dk, dv, (u64)cmax, 0, 64);
		cub::DoubleBuffer<u32> di((u32*)posA.p, (u32*)posB.p);
		cub::DeviceRadixSort::SortPairs(nullptr, c, dk, di, (u64)cmax, 0, 64);
		cub::DeviceScan::InclusiveScan(nullptr, d, headv.p, headv.p, MaxOp(), (u64)cmax);
		cub::DeviceSelect::Flagged(nullptr, e, posA.p, active.p, posB.p, d_count.p, (::cuda::std::int64_t)cmax);
		temp_bytes = std::max(std::max(a, b), std::max(c, std::max(d, e))) + 1024;
	}
	CKI(temp.get(temp_bytes));
	long long rounds_total = 0;
	for (const Range& rg : ranges)
	{
		if (rg.count == 0) continue;
		const u64 C = rg.count; u64* sa_b = SA.p + rg.base;
		size_t tb = temp_bytes;
		{
			thrust::counting_iterator<u64> it(0); InRange pr{T.p, rg.lo, rg.hi};
			CKI(cub::DeviceSelect::If(temp.p, tb, it, posA.p, d_count.p, (::cuda::std::int64_t)N, pr));
		}
		u64 got = 0; CKI(cudaMemcpy(&got, d_count.p, 8, cudaMemcpyDeviceToHost));
		if (got != C) { g_err = "kb_index_build: selection and histogram disagree"; return KB_ECUDA; }
		k_keys<<<grid_for(C), 256>>>(T.p, N, posA.p, C, 0, keyA.p);
		cub::DoubleBuffer<u64> dk(keyA.p, keyB.p), dv(posA.p, posB.p);
		tb = temp_bytes; CKI(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, C, 0, 64));
		k_first_round<<<grid_for(C), 256>>>(dk.Current(), dv.Current(), C, sa_b, headv.p, active.p);
		CKI(cudaGetLastError());
		// the tied suffixes: slot in the bucket, position, run id (= slot of the run's head)
		if (slotA.n < C) { CKI(slotA.get(C)); CKI(slotB.get(C)); CKI(gidA.get(C)); CKI(gidB.get(C)); CKI(idxA.get(C)); CKI(idxB.get(C)); }
		tb = temp_bytes; CKI(cub::DeviceScan::InclusiveScan(temp.p, tb, headv.p, gidB.p, MaxOp(), C));   // run id per slot
		u64* pos_sorted = dv.Current(); u64* pos_other = dv.Alternate();
		{
			thrust::counting_iterator<u32> slots(0);
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, slots, active.p, slotA.p, d_count.p, (::cuda::std::int64_t)C));
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, gidB.p, active.p, gidA.p, d_count.p, (::cuda::std::int64_t)C));
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, pos_sorted, active.p, pos_other, d_count.p, (::cuda::std::int64_t)C));
		}
		u64 n64 = 0; CKI(cudaMemcpy(&n64, d_count.p, 8, cudaMemcpyDeviceToHost));
		// Buffers of the refinement rounds. The tied suffixes, in ascending slot order: P (positions) stays in one position buffer,
		// G (run ids) in gidA, S (slots) alternates between slotA and slotB; everything else is scratch.
		u64* P = pos_other; u64* Pn = pos_sorted;
		u32* S = slotA.p; u32* Sn = slotB.p; u32* G = gidA.p; u32* Galt = gidB.p;
		u64 depth = 29; int round = 0;
		while (n64 > 0)
		{
			const u32 n = (u32)n64;
			u64* K = keyA.p;   // key of every tied suffix, indexable by its place in the list
			k_keys_idx<<<grid_for(n), 256>>>(T.p, N, P, n, depth, K, idxA.p);
			// order by (run id, key): sort a copy of the keys carrying the list index, then sort stably by run id
			CKI(cudaMemcpyAsync(keyB.p, K, (size_t)n * 8, cudaMemcpyDeviceToDevice, 0));
			cub::DoubleBuffer<u64> rk(keyB.p, Pn);   // Pn is free until k_next_round writes it
			cub::DoubleBuffer<u32> ri(idxA.p, idxB.p);
			tb = temp_bytes; CKI(cub::DeviceRadixSort::SortPairs(temp.p, tb, rk, ri, (u64)n, 0, 64));
			k_gather_u32<<<grid_for(n), 256>>>(G, ri.Current(), n, headv.p);   // run ids in key order
			int bits = 1; while (bits < 32 && (C >> bits) != 0) bits++;
			cub::DoubleBuffer<u32> gk(headv.p, Galt), gi(ri.Current(), ri.Alternate());
			tb = temp_bytes; CKI(cub::DeviceRadixSort::SortPairs(temp.p, tb, gk, gi, (u64)n, 0, bits));
			u32* gid_sorted = gk.Current(); u32* order = gi.Current(); u32* head_slot = gk.Alternate();
			k_next_round<<<grid_for(n), 256>>>(K, P, S, gid_sorted, order, n, sa_b, Pn, head_slot, active.p);
			CKI(cudaGetLastError());
			// new run ids (slot of the latest head), then keep the suffixes that are still tied
			u32* gid_new = gi.Alternate();   // `order` was read for the last time above; its twin is free
			tb = temp_bytes; CKI(cub::DeviceScan::InclusiveScan(temp.p, tb, head_slot, gid_new, MaxOp(), (u64)n));
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, S, active.p, Sn, d_count.p, (::cuda::std::int64_t)n));
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, Pn, active.p, P, d_count.p, (::cuda::std::int64_t)n));
			tb = temp_bytes; CKI(cub::DeviceSelect::Flagged(temp.p, tb, gid_new, active.p, G, d_count.p, (::cuda::std::int64_t)n));
			CKI(cudaMemcpy(&n64, d_count.p, 8, cudaMemcpyDeviceToHost));
			std::swap(S, Sn);
			depth += 29; round++;
			if (depth > N + 64) { g_err = "kb_index_build: refinement did not terminate"; return KB_ECUDA; }
		}
		rounds_total += round;
		if (verbose) { char b[96]; snprintf(b, sizeof(b), "bucket [%u,%u) %llu suffixes, %d rounds", rg.lo, rg.hi, (unsigned long long)C, round); lap(b); }
	}
	keyA.drop(); keyB.drop(); posA.drop(); posB.drop(); slotA.drop(); slotB.drop(); gidA.drop(); gidB.drop(); idxA.drop(); idxB.drop(); headv.drop(); active.drop();
	lap("suffix array");

	// ---- BWT with interleaved Occ counts, samples ----
	Buf<unsigned long long> d_primary; CKI(d_primary.get(1)); CKI(cudaMemset(d_primary.p, 0, 8));
	k_find_primary<<<grid_for(N), 256>>>(SA.p, N, d_primary.p);
	unsigned long long primary = 0; CKI(cudaMemcpy(&primary, d_primary.p, 8, cudaMemcpyDeviceToHost));
	if (primary == 0) { g_err = "kb_index_build: suffix 0 not found"; return KB_ECUDA; }
	const u64 nw = (N + 15) / 16, nblk = (N + 127) / 128, n_occ = nblk + 1, bwt_words = nw + n_occ * 8;
	const u64 tail_at = (N / 128) * 16 + ((N % 128) ? 8 + ((N % 128) + 15) / 16 : 0);
	Buf<u32> d_bwt, cnt4; Buf<u64> cA, cC, cG, cT;
	CKI(d_bwt.get(bwt_words)); CKI(cnt4.get(nw)); CKI(cA.get(nblk + 1)); CKI(cC.get(nblk + 1)); CKI(cG.get(nblk + 1)); CKI(cT.get(nblk + 1));
	CKI(cudaMemset(d_bwt.p, 0, bwt_words * 4));
	k_bwt_words<<<grid_for(nw), 256>>>(T.p, SA.p, N, primary, d_bwt.p, cnt4.p);
	CKI(cudaMemset(cA.p + nblk, 0, 8)); CKI(cudaMemset(cC.p + nblk, 0, 8)); CKI(cudaMemset(cG.p + nblk, 0, 8)); CKI(cudaMemset(cT.p + nblk, 0, 8));
	k_block_counts<<<grid_for(nblk), 256>>>(cnt4.p, nw, nblk, cA.p, cC.p, cG.p, cT.p);
	CKI(cudaGetLastError());
	for (Buf<u64>* c : {&cA, &cC, &cG, &cT})
	{
		size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, c->p, c->p, nblk + 1);
		if (tb > temp.n) CKI(temp.get(tb));
		CKI(cub::DeviceScan::ExclusiveSum(temp.p, tb, c->p, c->p, nblk + 1));
	}
	k_put_counts<<<grid_for(nblk + 1), 256>>>(cA.p, cC.p, cG.p, cT.p, nblk, tail_at, d_bwt.p);
	CKI(cudaGetLastError());
	u64 tot[4];
	CKI(cudaMemcpy(&tot[0], cA.p + nblk, 8, cudaMemcpyDeviceToHost)); CKI(cudaMemcpy(&tot[1], cC.p + nblk, 8, cudaMemcpyDeviceToHost));
	CKI(cudaMemcpy(&tot[2], cG.p + nblk, 8, cudaMemcpyDeviceToHost)); CKI(cudaMemcpy(&tot[3], cT.p + nblk, 8, cudaMemcpyDeviceToHost));
	const u64 n_sa = (N + 32) / 32;
	Buf<u64> smp; CKI(smp.get(n_sa));
	k_samples<<<grid_for(n_sa), 256>>>(SA.p, n_sa, smp.p);
	CKI(cudaGetLastError());
	lap("bwt + samples");
	out->primary = primary; out->seq_len = N; out->L2[0] = 0;
	for (int c = 0; c < 4; c++) out->L2[c + 1] = out->L2[c] + tot[c];
	if (out->L2[4] != N) { g_err = "kb_index_build: symbol counts do not add up to the text length"; return KB_ECUDA; }
	out->bwt_words = bwt_words; out->n_sa = n_sa;
	CKI(cudaMallocHost((void**)&out->bwt, bwt_words * 4));
	if (cudaMallocHost((void**)&out->sa, (n_sa > 1 ? n_sa - 1 : 1) * 8) != cudaSuccess) { cudaFreeHost(out->bwt); out->bwt = nullptr; return fail("cudaMallocHost(sa)", cudaErrorMemoryAllocation); }
	CKI(cudaMemcpy(out->bwt, d_bwt.p, bwt_words * 4, cudaMemcpyDeviceToHost));
	if (n_sa > 1) CKI(cudaMemcpy(out->sa, smp.p, (n_sa - 1) * 8, cudaMemcpyDeviceToHost));
	lap("copied back");
	if (verbose) fprintf(stderr, "[kb index] %llu suffixes, %zu super-buckets, %lld refinement rounds\n", (unsigned long long)N, ranges.size(), rounds_total);
	cudaEventDestroy(ev0); cudaEventDestroy(ev1);
	return KB_OK;
}

}   // extern "C"
