// kart_b200: CUDA kernels (sm_100a) and the C ABI declared in include/kart_b200.h.
// One context = one device, one stream. The pipeline of a batch is six kernels with no host round trip in between:
//   k_fm_seed -> k_sa_locate -> k_cand_pair -> k_rescue_{plan,win,commit} -> k_segments -> k_align -> k_assemble -> k_finalize
// Capacities of the bump-allocated arenas are checked on the device; a batch that overflowed anything is rerun with
// larger arenas (never silently truncated, never sent to a CPU path -- there is none).
#ifndef KB_EMUL
#include <cuda_runtime.h>
#define KB_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define KB_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#else   // host emulation build for the CPU-only test container (tests/emul/cuda_shim.h); never part of the product library
#define KB_LAUNCH(kern, grid, block, stream, ...) kb_emul_launch((grid), (block), [&]() { kern(__VA_ARGS__); })
#define KB_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) do { (void)(smem); kb_emul_launch((grid), (block), [&]() { kern(__VA_ARGS__); }); } while (0)
#endif
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>
#include <thread>
#include "../../include/kart_b200.h"
#include "kb_stages.cuh"

#define KB_BLOCK 128
#define KB_SLOTS 3           // batches in flight in kb_map_chunk's pipeline
#define KB_ALIGN_WARPS (148 * 8)   // warps of the warp-per-job kernels (k_align_part, k_nw_warp), each with a private arena
#define KB_ALIGN_POOL 10240  // shared-memory bytes per warp of k_align (fragment chars, codes, 2-bit traceback)

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void kb_warp_add64(unsigned long long* dst, unsigned long long v)
{
#ifndef KB_EMUL
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
	if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
#else
	*dst += v;
#endif
}

// read characters -> packed words (kb_fm.cuh "packed reads"); one thread per (read, word)
__global__ void __launch_bounds__(KB_BLOCK) k_pack(KbBatchDev bt)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (long long)bt.n_reads * bt.pk_wpr) kb_pack_word(bt, (int)(t / bt.pk_wpr), (int)(t % bt.pk_wpr));
}

// packed reads (kb_reads_packed_t) -> KbPk words and the read characters; one thread per (read, word). The characters that are
// not upper-case bases follow in k_unpack_exc.
#ifndef KB_EMUL
// The characters of a warp's 32 (read, word) pieces are one contiguous stretch of `seq` (reads lie back to back, pieces in order), but
// every piece starts 32 (or fewer) bytes after its neighbour at whatever alignment the read lengths give: written piece by piece that is
// 32 one-byte stores per thread, each touching 32 sectors per warp (ncu r21, C3: 0.82 ms per million reads, as long as k_segments; it is
// the difference between the end-to-end and the device-resident step). So the warp stages its stretch in shared memory at the same
// alignment modulo 16 and writes it out with 16-byte stores, lane after lane.
__global__ void __launch_bounds__(KB_BLOCK) k_unpack(KbBatchDev bt, const u64* code, u8* seq)
{
	__shared__ __align__(16) u8 stage[KB_BLOCK / 32][32 * 32 + 32];
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31; u8* st = stage[threadIdx.x >> 5];
	int n = 0; u64 dst = 0, c = 0;
	if (t < (long long)bt.n_reads * bt.pk_wpr)
	{
		const int r = (int)(t / bt.pk_wpr), w = (int)(t % bt.pk_wpr);
		const u64 off = bt.seq_off[r]; const int len = (int)(bt.seq_off[r + 1] - off);
		if (32 * w < len)
		{
			n = len - 32 * w < 32 ? len - 32 * w : 32;
			const u64 at = (off >> 5) - (bt.seq_off[0] >> 5) + (u64)r + (u64)w;   // slot-local word index
			c = code[at];
			KbPk k; k.code = c; k.n4 = n == 32 ? 0u : (~0u >> n); k.bad = k.n4;
			bt.pk[(off >> 5) + (u64)r + (u64)w] = k;
			dst = (off - bt.seq_off[0]) + 32 * (u64)w;
		}
	}
	const u32 have = __ballot_sync(0xFFFFFFFFu, n > 0);
	if (have == 0) return;
	const int first = __ffs(have) - 1, last = 31 - __clz(have);
	const u64 start = __shfl_sync(0xFFFFFFFFu, dst, first);
	const u64 end = __shfl_sync(0xFFFFFFFFu, dst + (u64)n, last);
	const u32 a0 = (u32)(start & 15u);                       // the stretch keeps its alignment modulo 16 in shared memory
	if (n > 0)
	{
		const u32 rel = a0 + (u32)(dst - start);
		u32 wd[8];
		for (int q = 0; q < 8; q++)
		{
			u32 v = 0;
			for (int i = 0; i < 4; i++) v |= ((0x54474341u >> (8 * (u32)((c >> (62 - 2 * (4 * q + i))) & 3ull))) & 0xFFu) << (8 * i);   // "ACGT"
			wd[q] = v;
		}
		if (n == 32 && (rel & 3u) == 0u) { u32* p = reinterpret_cast<u32*>(st + rel); for (int q = 0; q < 8; q++) p[q] = wd[q]; }
		else if (n == 32 && (rel & 1u) == 0u) { unsigned short* p = reinterpret_cast<unsigned short*>(st + rel); for (int q = 0; q < 8; q++) { p[2 * q] = (unsigned short)wd[q]; p[2 * q + 1] = (unsigned short)(wd[q] >> 16); } }
		else for (int i = 0; i < n; i++) st[rel + i] = (u8)(wd[i >> 2] >> (8 * (i & 3)));
	}
	__syncwarp();
	const u32 total = (u32)(end - start);
	u8* out = seq + start - a0;                              // 16-byte aligned (the slot's buffer is)
	const u32 lo = a0, hi = a0 + total;                      // the bytes of `st` that hold characters
	const u32 body_lo = (lo + 15u) & ~15u, body_hi = hi & ~15u;
	if (body_lo < body_hi)
	{
		for (u32 b = body_lo + 16u * (u32)lane; b < body_hi; b += 16u * 32u) *reinterpret_cast<uint4*>(out + b) = *reinterpret_cast<const uint4*>(st + b);
		for (u32 b = lo + (u32)lane; b < body_lo; b += 32u) out[b] = st[b];
		for (u32 b = body_hi + (u32)lane; b < hi; b += 32u) out[b] = st[b];
	}
	else for (u32 b = lo + (u32)lane; b < hi; b += 32u) out[b] = st[b];
}
#else
__global__ void __launch_bounds__(KB_BLOCK) k_unpack(KbBatchDev bt, const u64* code, u8* seq)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)bt.n_reads * bt.pk_wpr) return;
	const int r = (int)(t / bt.pk_wpr), w = (int)(t % bt.pk_wpr);
	const u64 off = bt.seq_off[r]; const int len = (int)(bt.seq_off[r + 1] - off);
	if (32 * w >= len) return;
	const int n = len - 32 * w < 32 ? len - 32 * w : 32;
	const u64 at = (off >> 5) - (bt.seq_off[0] >> 5) + (u64)r + (u64)w;   // slot-local word index
	const u64 c = code[at];
	KbPk k; k.code = c; k.n4 = n == 32 ? 0u : (~0u >> n); k.bad = k.n4;
	bt.pk[(off >> 5) + (u64)r + (u64)w] = k;
	u8* s = seq + (off - bt.seq_off[0]) + 32 * w;
	for (int i = 0; i < n; i++) s[i] = (u8)(0x54474341u >> (8 * (u32)((c >> (62 - 2 * i)) & 3ull)));   // "ACGT"
}
#endif
__global__ void __launch_bounds__(KB_BLOCK) k_unpack_exc(KbBatchDev bt, const u64* exc, u32 n_exc, u32 first_read, u8* seq)
{
	const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_exc) return;
	const u64 e = exc[t]; const u32 r = (u32)(e >> 32) - first_read, pos = (u32)(e >> 8) & 0xFFFFFFu; const u8 ch = (u8)e;
	const u64 off = bt.seq_off[r];
	seq[(off - bt.seq_off[0]) + pos] = ch;
	u32* f = reinterpret_cast<u32*>(&bt.pk[(off >> 5) + (u64)r + (u64)(pos >> 5)]) + 2;   // KbPk: {u64 code; u32 n4; u32 bad}
	const u32 bit = 1u << (31 - (pos & 31));
	if (kb_nt4(ch) > 3) KB_ATOMIC_OR(f, bit);
	KB_ATOMIC_OR(f + 1, bit);
}

// .pac bytes -> big-endian 64-bit words (upload time only)
__global__ void k_ref64(const u8* pac, u64 bytes, u64 words, u64* out)
{
	u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= words) return;
	u64 v = 0;
	for (int i = 0; i < 8; i++) { u64 b = w * 8 + i; v = (v << 8) | (u64)(b < bytes ? pac[b] : 0); }
	out[w] = v;
}

// MINB = resident blocks per SM the register allocation is tuned for: 10 (48 registers, more loads in flight: HBM-sized indexes)
// or 8 (64 registers, no spills: L2-resident indexes, where the kernel is ALU-bound)
// ROW = u32 when every BWT row number fits 32 bits (kb_extend), u64 for the multi-Gbp indexes
template <int MINB, class ROW>
__global__ void __launch_bounds__(KB_BLOCK, MINB) k_fm_seed(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	u32 steps = 0, blocks = 0;
	kb_seed_read<ROW>(ix, pm, bt, r, r < bt.n_reads, &steps, &blocks);
	kb_warp_add64(&bt.work[0], steps); kb_warp_add64(&bt.work[1], blocks);
}

// lane-queue seeding (kb_seed_lane): a fixed grid of warps, each with a contiguous range of reads that its lanes draw from
#ifndef KB_EMUL
// reads come from the warp's range through a shared-memory counter; seed slices are handed out from a shared-memory cursor
// relative to the warp's slab, which gets its place in the arena with ONE global atomic when the warp is done (a global
// atomic per read, from two or three lanes at a time, was 12 % of the kernel's stall samples: ncu r15)
struct KbSeedWarpQ
{
	u32* cnt; u32 end; u32* seeds; u32 max_ns; KbPk* stg;
	__device__ __forceinline__ KbPk* stage() { return stg; }
	__device__ __forceinline__ int next() { const u32 k = atomicAdd(cnt, 1u); return k < end ? (int)k : -1; }
	__device__ __forceinline__ void store(const KbBatchDev& bt, int rd, int ns) { bt.seed_off[rd] = atomicAdd(seeds, (u32)ns); if ((u32)ns > max_ns) max_ns = (u32)ns; }
};
#endif
template <int MINB, class ROW>
__global__ void __launch_bounds__(KB_BLOCK, MINB) k_fm_seed_q(KbIndexDev ix, KbParams pm, KbBatchDev bt, int qp, int qs, int trips, int tail_max, int staged)
{
	u32 steps = 0, blocks = 0;
#ifndef KB_EMUL
	extern __shared__ __align__(16) u8 seed_stage[];   // KB_SEED_STAGE packed words per lane when `staged` (see kb_fm.cuh)
	__shared__ u32 cnt[KB_BLOCK / 32], wseeds[KB_BLOCK / 32];
	const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5, wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 lo = (u32)((u64)bt.n_reads * gwarp / nwarps), hi = (u32)((u64)bt.n_reads * (gwarp + 1) / nwarps);
	if (lane == 0) { cnt[wib] = lo; wseeds[wib] = 0; }
	__syncwarp();
	KbSeedWarpQ q; q.cnt = &cnt[wib]; q.end = hi; q.seeds = &wseeds[wib]; q.max_ns = 0;
	q.stg = staged ? reinterpret_cast<KbPk*>(seed_stage) + (size_t)threadIdx.x * KB_SEED_STAGE : nullptr;
	kb_seed_lane<ROW>(ix, pm, bt, q, &steps, &blocks, qp, qs, trips, tail_max);
	__syncwarp();
	{
		const u32 total = wseeds[wib]; u32 base = 0;
		if (lane == 0 && total) base = atomicAdd(&bt.counters[0], total);
		base = __shfl_sync(0xFFFFFFFFu, base, 0);
		if (total) for (u32 r = lo + lane; r < hi; r += 32) bt.seed_off[r] += base;
		if (lane == 0 && (u64)base + (u64)total > (u64)bt.cap_segs) atomicOr(&bt.counters[3], (u32)KB_OVF_SEEDS);
		u32 m = q.max_ns;
		for (int o = 16; o > 0; o >>= 1) { const u32 v = __shfl_xor_sync(0xFFFFFFFFu, m, o); if (v > m) m = v; }
		if (lane == 0 && m) atomicMax(&bt.counters[5], m);
	}
#else
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < bt.n_reads; r += gridDim.x * blockDim.x) { KbSeedOne q; q.r = r; q.staged = staged != 0; kb_seed_lane<ROW>(ix, pm, bt, q, &steps, &blocks, qp, qs, trips, tail_max); }
#endif
	kb_warp_add64(&bt.work[0], steps); kb_warp_add64(&bt.work[1], blocks);
}

__global__ void __launch_bounds__(KB_BLOCK) k_sa_locate(KbIndexDev ix, KbBatchDev bt)
{
	long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	u32 lf = 0;
	if (t < (long long)bt.n_reads * bt.max_hits) kb_locate_hit(ix, bt, (int)(t / bt.max_hits), (int)(t % bt.max_hits), &lf);
	kb_warp_add64(&bt.work[2], lf);
}
// with the full SA a locate is one load: a thread per (read, search) slot mostly finds nothing to do (3.3 seeds per read against
// 12 slots), so one thread takes all the searches of its read
__global__ void __launch_bounds__(KB_BLOCK) k_sa_locate_reads(KbIndexDev ix, KbBatchDev bt)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	u32 lf = 0;
	if (r < bt.n_reads) { const int nh = bt.n_hits[r]; for (int h = 0; h < nh; h++) kb_locate_hit(ix, bt, r, h, &lf); }
	kb_warp_add64(&bt.work[2], lf);
}

// expands the sampled SA into a full one (upload time only)
__global__ void k_expand_sa(KbIndexDev ix, u64* full)
{
	u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (k > ix.seq_len) return;
	u32 steps; full[k] = kb_sa(ix, k, &steps);
}

// seeding table (upload time only)
__global__ void k_build_ktab(KbIndexDev ix, int K, KbKtab* out)
{
	u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < (1ull << (2 * K))) out[t] = kb_ktab_entry(ix, t, K);
}

// re-blocks the BWA Occ/BWT interleave (16 words / 128 rows, u64 counts) into 8 words / 64 rows: u32 counts + two bit planes
__global__ void k_reblock(const u32* bwt, u64 bwt_words, u64 n_new, u32* occ)
{
	u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_new) return;
	u64 ob = (j >> 1) * 16; int half = (int)(j & 1);
	u32 cnt[4], w[8];
	for (int c = 0; c < 4; c++) { u64 i = ob + 2 * c; cnt[c] = i < bwt_words ? bwt[i] : 0; }   // low words of the u64 counts (checked < 2^32 on the host)
	for (int k = 0; k < 8; k++) { u64 i = ob + 8 + k; w[k] = i < bwt_words ? bwt[i] : 0; }
	if (half)
	{
		kb_count32(((u64)w[0] << 32) | w[1], 32, cnt);
		kb_count32(((u64)w[2] << 32) | w[3], 32, cnt);
	}
	u64 lo = 0, hi = 0;
	for (int i = 0; i < 64; i++)
	{
		u32 code = (w[4 * half + (i >> 4)] >> ((~i & 15) << 1)) & 3u;
		lo |= (u64)(code & 1u) << (63 - i); hi |= (u64)(code >> 1) << (63 - i);
	}
	u32* o = occ + j * 8;
	o[0] = cnt[0]; o[1] = cnt[1]; o[2] = cnt[2]; o[3] = cnt[3];
	o[4] = (u32)lo; o[5] = (u32)(lo >> 32); o[6] = (u32)hi; o[7] = (u32)(hi >> 32);
}

__global__ void __launch_bounds__(KB_BLOCK) k_cand_pair(KbIndexDev ix, KbParams pm, KbBatchDev bt, int split_heavy) { kb_stage_cand_pair(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x, split_heavy != 0); }
// the items k_cand_pair put on the heavy list (a read with more than KB_CAND_HEAVY seeds): one warp each
#ifndef KB_EMUL
__global__ void __launch_bounds__(KB_BLOCK) k_cand_heavy(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	__shared__ KbWarpSort sw[KB_BLOCK / 32];
	if (bt.counters[3]) return;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	KbWarpSort& w = sw[wib];
	const u32 count = bt.counters[15];
	for (u32 q = gwarp; q < count; q += nwarps)
	{
		const int t = bt.slow_list2[q];
		for (int which = 0; which < 2; which++)
		{
			KbSeg* v; KbSeg* tmp; int n;
			if (!kb_cand_heavy_list(pm, bt, t, which, &v, &n, &tmp)) break;
			if (n < 2) continue;
			if (n > KB_WSORT_MAX) { if (lane == 0) kb_sort_segs<false>(v, n); __syncwarp(); continue; }
			if (lane == 0) kb_wsort_begin(w, v, n, tmp);
			__syncwarp();
			kb_wsort_load(w, lane); __syncwarp();
			for (int k = 2; k <= w.p; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) { kb_wsort_step(w, k, j, lane); __syncwarp(); }
			kb_wsort_gather(w, lane); __syncwarp();
			kb_wsort_scatter(w, lane); __syncwarp();
		}
	}
}
#else
static void k_cand_heavy(KbIndexDev ix, KbParams pm, KbBatchDev bt)   // emulation: a warp = a loop over 32 lanes per phase
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	static thread_local KbWarpSort w;
	const u32 count = bt.counters[15];
	for (u32 q = 0; q < count; q++)
	{
		const int t = bt.slow_list2[q];
		for (int which = 0; which < 2; which++)
		{
			KbSeg* v; KbSeg* tmp; int n;
			if (!kb_cand_heavy_list(pm, bt, t, which, &v, &n, &tmp)) break;
			if (n < 2) continue;
			if (n > KB_WSORT_MAX) { kb_sort_segs<false>(v, n); continue; }
			kb_wsort_begin(w, v, n, tmp);
			for (int l = 31; l >= 0; l--) kb_wsort_load(w, l);
			for (int k = 2; k <= w.p; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) for (int l = 31; l >= 0; l--) kb_wsort_step(w, k, j, l);
			for (int l = 31; l >= 0; l--) kb_wsort_gather(w, l);
			for (int l = 31; l >= 0; l--) kb_wsort_scatter(w, l);
		}
	}
}
#endif
// ... and then a thread each for the scan over the sorted seeds, the pairing and the pruning. On the sorting warp's lane 0 that part
// ran one item after the other (80 % of k_cand_heavy's samples at one lane: ncu r21, C3); here every heavy item of the batch is in
// flight at once, and a warp lasts as long as its longest item instead of as long as the sum of thirteen.
__global__ void __launch_bounds__(KB_BLOCK) k_cand_heavy_finish(KbIndexDev ix, KbParams pm, KbBatchDev bt, int wide)
{
	if (bt.counters[3]) return;
	const u32 count = bt.counters[15];
	if (wide) { for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) kb_cand_finish<true>(ix, pm, bt, bt.slow_list2[q]); }
	else for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) kb_cand_finish<false>(ix, pm, bt, bt.slow_list2[q]);
}
__global__ void __launch_bounds__(KB_BLOCK) k_cand_pacbio(KbIndexDev ix, KbParams pm, KbBatchDev bt, int sorted) { kb_stage_cand_pacbio(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, sorted != 0); }
// A 7-kbp read brings ~600 seeds; sorted by the read's own thread (Shell sort over 24-byte records in HBM) that was 79 % of k_cand_pacbio's
// samples (ncu r24, C5). A warp per read sorts them first, with the same bitonic network on packed keys as k_cand_heavy; the read's
// unused candidate slots (n_seeds + 1 of them, as large as a seed) are the permutation buffer. Lists beyond KB_WSORT_MAX seeds, and reads of
// a megabase or more (rPos would not fit the key), are sorted by one lane as before.
#ifndef KB_EMUL
__global__ void __launch_bounds__(KB_BLOCK) k_cand_pacbio_sort(KbBatchDev bt)
{
	__shared__ KbWarpSort sw[KB_BLOCK / 32];
	if (bt.counters[3]) return;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	KbWarpSort& w = sw[wib];
	for (u32 r = gwarp; r < (u32)bt.n_reads; r += nwarps)
	{
		const int n = bt.n_seeds[r];
		if (n < 2) continue;
		KbSeg* v = bt.segs + bt.seed_off[r];
		if (n > KB_WSORT_MAX || bt.seq_off[r + 1] - bt.seq_off[r] >= (1ull << 20)) { if (lane == 0) kb_sort_segs<true>(v, n); __syncwarp(); continue; }
		if (lane == 0) kb_wsort_begin(w, v, n, reinterpret_cast<KbSeg*>(bt.cands + bt.cand_off[r]), true);
		__syncwarp();
		kb_wsort_load(w, lane); __syncwarp();
		for (int k = 2; k <= w.p; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) { kb_wsort_step(w, k, j, lane); __syncwarp(); }
		kb_wsort_gather(w, lane); __syncwarp();
		kb_wsort_scatter(w, lane); __syncwarp();
	}
}
#else
static void k_cand_pacbio_sort(KbBatchDev bt)   // emulation: a warp = a loop over 32 lanes per phase
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	static thread_local KbWarpSort w;
	for (int r = 0; r < bt.n_reads; r++)
	{
		const int n = bt.n_seeds[r];
		if (n < 2) continue;
		KbSeg* v = bt.segs + bt.seed_off[r];
		if (n > KB_WSORT_MAX || bt.seq_off[r + 1] - bt.seq_off[r] >= (1ull << 20)) { kb_sort_segs<true>(v, n); continue; }
		kb_wsort_begin(w, v, n, reinterpret_cast<KbSeg*>(bt.cands + bt.cand_off[r]), true);
		for (int l = 31; l >= 0; l--) kb_wsort_load(w, l);
		for (int k = 2; k <= w.p; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) for (int l = 31; l >= 0; l--) kb_wsort_step(w, k, j, l);
		for (int l = 31; l >= 0; l--) kb_wsort_gather(w, l);
		for (int l = 31; l >= 0; l--) kb_wsort_scatter(w, l);
	}
}
#endif
// rescue: plan (thread per job) -> windows (block per task) -> commit (thread per job); see kb_pair.cuh "task-parallel rescue"
__global__ void __launch_bounds__(KB_BLOCK) k_rescue_plan(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	if (bt.counters[3]) return;
	const int count = (int)bt.counters[4];
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) kb_rescue_plan(ix, pm, bt, k);
}
__global__ void __launch_bounds__(KB_BLOCK) k_rescue_commit(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	if (bt.counters[3]) return;
	const int count = (int)bt.counters[4];
	u32 attempted = 0;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) kb_rescue_commit(pm, bt, k, &attempted);
#ifndef KB_EMUL
	for (int o = 16; o > 0; o >>= 1) attempted += __shfl_down_sync(0xFFFFFFFFu, attempted, o);
	if ((threadIdx.x & 31) != 0) attempted = 0;
#endif
	if (attempted) KB_ATOMIC_ADD(&bt.counters[7], attempted);
}
// KB_RESCUE_FAST=0: every window on the slow list
__global__ void __launch_bounds__(KB_BLOCK) k_rescue_all_slow(KbBatchDev bt)
{
	if (bt.counters[3]) return;
	const u32 count = bt.counters[27];
	for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) bt.rslow[q] = q;
	if (blockIdx.x == 0 && threadIdx.x == 0) bt.counters[29] = count;
}
// warp per rescue window out of shared memory (kb_pair.cuh "warp-per-window fast path"); what it cannot take goes to the slow list
#ifndef KB_EMUL
__global__ void __launch_bounds__(KB_BLOCK) k_rescue_fast(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	extern __shared__ __align__(16) u8 rf_pool[];
	__shared__ u32 ticket[KB_BLOCK / 32];
	if (bt.counters[3]) return;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	KbRescueFast& w = reinterpret_cast<KbRescueFast*>(rf_pool)[wib];
	const u32 count = bt.counters[27], batch = (u32)(bt.rf_batch > 0 ? bt.rf_batch : 1);
	if (lane == 0) w.have_mate = -1;
	while (true)
	{
		// a warp draws `batch` consecutive windows: a job's windows follow each other, so most of them face the mate whose index is in place
		if (lane == 0) ticket[wib] = atomicAdd(&bt.counters[25], batch);
		__syncwarp();
		const u32 q0 = ticket[wib];
		__syncwarp();
		if (q0 >= count) break;
		const u32 q1 = q0 + batch < count ? q0 + batch : count;
		for (u32 q = q0; q < q1; q++)
		{
			if (lane == 0) kb_rf_begin(ix, bt, w, q);
			__syncwarp();
			kb_rf_load(ix, bt, w, lane); __syncwarp();
			kb_rf_fill(w, lane); __syncwarp();
			kb_rf_scan(w, lane); __syncwarp();
			kb_rf_pairs(w, lane); __syncwarp();
			if (lane == 0) kb_rf_end(pm, bt, w);
			__syncwarp();
		}
	}
}
#else
static void k_rescue_fast(KbIndexDev ix, KbParams pm, KbBatchDev bt)   // emulation: a warp = a loop over 32 lanes per phase
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	static thread_local KbRescueFast w;
	const u32 count = bt.counters[27];
	w.have_mate = -1;
	for (u32 q = 0; q < count; q++)
	{
		kb_rf_begin(ix, bt, w, q);
		for (int t = 31; t >= 0; t--) kb_rf_load(ix, bt, w, t);
		for (int t = 31; t >= 0; t--) kb_rf_fill(w, t);
		for (int t = 31; t >= 0; t--) kb_rf_scan(w, t);
		for (int t = 31; t >= 0; t--) kb_rf_pairs(w, t);   // reversed on purpose: the result must not depend on append order
		kb_rf_end(pm, bt, w);
	}
}
#endif
#ifndef KB_EMUL
__global__ void __launch_bounds__(KB_BLOCK) k_rescue_win(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	__shared__ KbRescueJob sj; __shared__ u32 ticket;
	if (bt.counters[3]) return;
	const u32 count = bt.counters[29]; const int tid = threadIdx.x, nth = blockDim.x;   // the windows k_rescue_fast handed back
	KbRescueJob* j = &sj; KbArena ar = kb_job_arena(bt, blockIdx.x, blockDim.x);
	while (true)
	{
		if (tid == 0) ticket = atomicAdd(&bt.counters[28], 1u);
		__syncthreads();
		if (ticket >= count) break;
		const u32 q = bt.rslow[ticket];
		if (tid == 0) kb_rt_begin(ix, bt, j, ar, bt.rtasks[q]);
		__syncthreads();
		if (!j->ovf)
		{
			kb_rj_window(ix, j, tid, nth); kb_rj_index_clear(bt, j, tid, nth); __syncthreads();
			kb_rj_ids(j, tid, nth); kb_rj_index_fill(j, tid, nth); __syncthreads();
			kb_rj_pairs(ix, j, tid, nth); __syncthreads();
		}
		if (tid == 0) kb_rt_end(pm, bt, j, &bt.rtasks[q]);
		__syncthreads();
	}
}
#else
static void k_rescue_win(KbIndexDev ix, KbParams pm, KbBatchDev bt)   // emulation: the same phases, barriers replaced by loops over tid
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	const u32 count = bt.counters[29]; const int nth = (int)blockDim.x;
	static thread_local KbRescueJob job; KbRescueJob* j = &job; KbArena ar = kb_job_arena(bt, 0, blockDim.x);
	for (u32 qi = 0; qi < count; qi++)
	{
		const u32 q = bt.rslow[qi];
		kb_rt_begin(ix, bt, j, ar, bt.rtasks[q]);
		if (!j->ovf)
		{
			for (int t = 0; t < nth; t++) { kb_rj_window(ix, j, t, nth); kb_rj_index_clear(bt, j, t, nth); }
			for (int t = nth - 1; t >= 0; t--) { kb_rj_ids(j, t, nth); kb_rj_index_fill(j, t, nth); }
			for (int t = nth - 1; t >= 0; t--) kb_rj_pairs(ix, j, t, nth);   // reversed on purpose: the result must not depend on append order
		}
		kb_rt_end(pm, bt, j, &bt.rtasks[q]);
	}
}
#endif
__global__ void __launch_bounds__(KB_BLOCK) k_segments(KbIndexDev ix, KbParams pm, KbBatchDev bt, int slab)
{
#ifndef KB_EMUL
	__shared__ u32 range[KB_BLOCK / 32][2];
	if (slab > 0)   // kb_alloc_segx
	{
		const int wib = threadIdx.x >> 5;
		if ((threadIdx.x & 31) == 0) { const u32 b = atomicAdd(&bt.counters[8], (u32)slab); range[wib][0] = b; range[wib][1] = b + (u32)slab; }
		__syncwarp();
		bt.segx_slab = &range[wib][0];
	}
#else
	(void)slab;
#endif
	kb_stage_segments(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(KB_BLOCK) k_segments_slow(KbIndexDev ix, KbParams pm, KbBatchDev bt) { kb_stage_segments_slow(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
// phase B (kb_align.cuh "phase B"): partition -> nw_alignment problems by size class -> gather
#ifndef KB_EMUL
// One warp per partition job. A job is a chain of short phases with global-memory latency in between (ALU pipe ~11 %, ncu
// r13), so throughput is the number of jobs in flight: the per-warp shared-memory pool is sized at launch (pool_bytes; what does
// not fit spills to the warp's HBM arena) to trade pool size against resident warps, and warps draw jobs from a ticket counter
// so that no SM is left holding a static share of long jobs.
__global__ void __launch_bounds__(KB_BLOCK) k_align_part(KbIndexDev ix, KbParams pm, KbBatchDev bt, int pool_bytes)
{
	__shared__ KbPartWarp sw[KB_BLOCK / 32];
	__shared__ u32 ticket[KB_BLOCK / 32];
	extern __shared__ __align__(16) u8 dyn_pool[];
	if (bt.counters[3]) return;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if ((int)gwarp >= bt.wscratch_warps) return;
	KbPartWarp& w = sw[wib];
	const u32 njobs = bt.counters[23];
	if (lane == 0) { w.ar.base = bt.wscratch + (u64)gwarp * bt.wscratch_per_warp; w.ar.cap = bt.wscratch_per_warp; w.fast.base = dyn_pool + (size_t)wib * (size_t)pool_bytes; w.fast.cap = (u64)pool_bytes; w.fast.ovf = false; }
	__syncwarp();
	while (true)
	{
		if (lane == 0) ticket[wib] = atomicAdd(&bt.counters[26], 1u);
		__syncwarp();
		const u32 q = ticket[wib];
		__syncwarp();
		if (q >= njobs) break;
		if (lane == 0) kb_pt_begin(ix, pm, bt, w, bt.part_list[q]);
		__syncwarp();
		kb_pt_fetch(ix, bt, w, lane);
		__syncwarp();
		while (w.ok)
		{
			if (lane == 0) w.state = w.it.next();
			__syncwarp();
			if (w.state != 2) break;
			w.it.part_scan(lane); __syncwarp();
			w.it.part_ids(lane); __syncwarp();
			w.it.part_pairs(lane); __syncwarp();
			if (lane == 0) w.state = w.it.part_grow() ? 3 : 2;
			__syncwarp();
			if (w.state == 3) { w.it.part_pairs(lane); __syncwarp(); }
			if (lane == 0) w.it.part_finish();
			__syncwarp();
		}
		if (lane == 0) kb_pt_end(bt, w);
		__syncwarp();
	}
}
__global__ void __launch_bounds__(KB_BLOCK) k_nw_warp(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	__shared__ KbPieceWarp sw[KB_BLOCK / 32];
	__shared__ __align__(16) u8 pool[KB_BLOCK / 32][KB_ALIGN_POOL];
	if (bt.counters[3]) return;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	if ((int)gwarp >= bt.wscratch_warps) return;
	KbPieceWarp& w = sw[wib];
	if (lane == 0) { w.ar.base = bt.wscratch + (u64)gwarp * bt.wscratch_per_warp; w.ar.cap = bt.wscratch_per_warp; w.fast.base = pool[wib]; w.fast.cap = KB_ALIGN_POOL; w.fast.ovf = false; w.cells = 0; w.calls = 0; }
	__syncwarp();
	// the wavefront class, plus a column-tile class (33..64, 65..128) when it holds too few problems to fill the machine with
	// one thread each: such a launch lasts as long as one thread's DP over a whole problem (0.2-0.3 ms, ncu r13), a warp does
	// the same problem in microseconds. k_nw_tile<4|5> steps aside on the same test.
	for (int cls = 4; cls < KB_NW_CLASSES; cls++)
	{
	const u32 count = bt.counters[16 + cls]; const u32* list = bt.piece_list + (size_t)cls * bt.cap_pieces;
	if (cls < KB_NW_CLASSES - 1 && count >= (u32)bt.nw_warp_below) continue;
	for (u32 q = gwarp; q < count; q += nwarps)
	{
		if (lane == 0) kb_pw_begin(bt, w, list[q]);
		__syncwarp();
		kb_pw_fetch(ix, w, lane);
		__syncwarp();
		if (w.ok)
		{
			kb_nww_init_rows(w.nw, lane);
			__syncwarp();
			while (true)
			{
				KbNwLane L;
				kb_nww_strip_begin(w.nw, L, lane);
				__syncwarp();
				const int steps = w.nw.n + w.nw.h - 1;
				for (int d = 0; d < steps; d++) { kb_nww_step(w.nw, L, d, lane); __syncwarp(); }
				if (lane == 0) w.more_strips = kb_nww_strip_end(w.nw) ? 1 : 0;
				__syncwarp();
				if (!w.more_strips) break;
			}
		}
		if (lane == 0) kb_pw_end(bt, w);
		__syncwarp();
	}
	}
	if (lane == 0) { if (w.cells) atomicAdd(&bt.work[3], w.cells); if (w.calls) atomicAdd(&bt.work[4], (unsigned long long)w.calls); }
}
#else
static void k_align_part(KbIndexDev ix, KbParams pm, KbBatchDev bt, int)   // emulation: the same phases, a warp = a loop over 32 lanes
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	static thread_local KbPartWarp w; static thread_local u8 pool[KB_ALIGN_POOL];
	w.ar.base = bt.wscratch; w.ar.cap = bt.wscratch_per_warp; w.fast.base = pool; w.fast.cap = KB_ALIGN_POOL; w.fast.ovf = false;
	const u32 njobs = bt.counters[23];
	for (u32 q = 0; q < njobs; q++)
	{
		kb_pt_begin(ix, pm, bt, w, bt.part_list[q]);
		for (int t = 31; t >= 0; t--) kb_pt_fetch(ix, bt, w, t);
		while (w.ok)
		{
			w.state = w.it.next();
			if (w.state != 2) break;
			for (int t = 0; t < 32; t++) w.it.part_scan(t);
			for (int t = 31; t >= 0; t--) w.it.part_ids(t);
			for (int t = 31; t >= 0; t--) w.it.part_pairs(t);
			if (w.it.part_grow()) for (int t = 31; t >= 0; t--) w.it.part_pairs(t);
			w.it.part_finish();
		}
		kb_pt_end(bt, w);
	}
}
static void k_nw_warp(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (bt.counters[3]) return;
	static thread_local KbPieceWarp w; KbNwLane L[32]; static thread_local u8 pool[KB_ALIGN_POOL];
	w.ar.base = bt.wscratch; w.ar.cap = bt.wscratch_per_warp; w.fast.base = pool; w.fast.cap = KB_ALIGN_POOL; w.fast.ovf = false; w.cells = 0; w.calls = 0;
	for (int cls = 4; cls < KB_NW_CLASSES; cls++)
	{
	const u32 count = bt.counters[16 + cls]; const u32* list = bt.piece_list + (size_t)cls * bt.cap_pieces;
	if (cls < KB_NW_CLASSES - 1 && count >= (u32)bt.nw_warp_below) continue;
	for (u32 q = 0; q < count; q++)
	{
		kb_pw_begin(bt, w, list[q]);
		for (int t = 31; t >= 0; t--) kb_pw_fetch(ix, w, t);
		if (w.ok)
		{
			for (int t = 0; t < 32; t++) kb_nww_init_rows(w.nw, t);
			while (true)
			{
				for (int t = 0; t < 32; t++) kb_nww_strip_begin(w.nw, L[t], t);
				const int steps = w.nw.n + w.nw.h - 1;
				for (int d = 0; d < steps; d++) for (int t = 31; t >= 0; t--) kb_nww_step(w.nw, L[t], d, t);   // lane order must not matter
				w.more_strips = kb_nww_strip_end(w.nw) ? 1 : 0;
				if (!w.more_strips) break;
			}
		}
		kb_pw_end(bt, w);
	}
	}
	bt.work[3] += w.cells; bt.work[4] += w.calls;
}
#endif
// one thread per nw_alignment problem of size class CLS: TW columns in registers, up to MAXM rows, MAXT column tiles
template <int CLS, int TW, int MAXM, int MAXT>
__global__ void __launch_bounds__(KB_BLOCK) k_nw_tile(KbIndexDev ix, KbParams pm, KbBatchDev bt)
{
	if (bt.counters[3]) return;
	if (CLS >= 4 && bt.counters[16 + CLS] < (u32)bt.nw_warp_below) return;   // k_nw_warp takes a sparse class
	unsigned long long cells = 0, calls = 0;
	kb_nwt_class<TW, MAXM, MAXT>(ix, bt, CLS, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, &cells, &calls);
	kb_warp_add64(&bt.work[3], cells); kb_warp_add64(&bt.work[4], calls);
}
__global__ void __launch_bounds__(KB_BLOCK) k_align_gather(KbBatchDev bt)
{
	// cigar elements are only handed out by the assemble kernels that follow: the cursor as it stands now and as k_finalize
	// finds it bracket this batch's elements in a chunk-wide arena (map_chunk_pipelined copies that range back on its own)
	if (blockIdx.x == 0 && threadIdx.x == 0) bt.counters[30] = KB_ATOMIC_ADD(bt.cig_cursor, 0u);
	if (bt.counters[3]) return;
	const u32 njobs = bt.counters[23];
	for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < njobs; q += gridDim.x * blockDim.x) kb_gather_job(bt, bt.part_list[q]);
}
__global__ void __launch_bounds__(KB_BLOCK) k_assemble(KbIndexDev ix, KbParams pm, KbBatchDev bt) { kb_stage_assemble(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x); }
__global__ void __launch_bounds__(KB_BLOCK) k_assemble_slow(KbIndexDev ix, KbParams pm, KbBatchDev bt) { kb_stage_assemble_slow(ix, pm, bt, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
// (r23: assembling a block's records in shared memory and copying them out in 16-byte words made this kernel slower, 0.98 -> 1.27 ms per
// 2.5 M reads: it is bound by the latency of its scattered report reads, and the barrier in front of the copy-out makes every warp wait
// for the block's slowest thread. The records are written straight from the thread again.)
__global__ void __launch_bounds__(KB_BLOCK) k_finalize(KbIndexDev ix, KbParams pm, KbBatchDev bt, kb_aln_t* aln)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) bt.counters[31] = KB_ATOMIC_ADD(bt.cig_cursor, 0u);
	const int t = blockIdx.x * blockDim.x + threadIdx.x, per = pm.paired ? 2 : 1, items = pm.paired ? (bt.n_reads >> 1) : bt.n_reads;
	if (t < items) kb_stage_finalize(ix, pm, bt, aln + (size_t)per * (size_t)t, t);
}

// ---- stage-level test entry (kb_debug_align): caller-chosen fragment pairs through classification, phase B and the per-segment
// part of phase C, so that the golden nw_alignment / fragment vectors and the oracle's Process*SequencePair reach
// k_align_part, k_nw_tile<*>, k_nw_warp and k_align_gather directly
__global__ void k_debug_classify(KbIndexDev ix, KbParams pm, KbBatchDev bt, const kb_dbg_frag_t* specs, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const kb_dbg_frag_t f = specs[i];
	const int r = (int)f.read;
	KbSeg sp; sp.gpos = f.gpos; sp.rpos = f.rpos; sp.rlen = f.rlen; sp.glen = f.glen; sp.simple = 0;
	const u8* seq = bt.seq + bt.seq_off[r]; const KbPk* rd = kb_pk_read(bt, r);
	KbSegX* out = &bt.segx[i];
	if (f.mode >= 3) { out->s = sp; out->info = KB_SEG_SKIP; out->aux = 0; kb_make_job(ix, bt, r, sp, f.mode == 4, out); }
	else kb_classify_segment(ix, pm, bt, r, seq, rd, sp, f.mode == 1 ? 0 : 1, f.mode == 0 ? 3 : 2, out);
}
__global__ void k_debug_assemble(KbIndexDev ix, KbParams pm, KbBatchDev bt, const kb_dbg_frag_t* specs, int n, kb_dbg_frag_out_t* res, u32* ops)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const kb_dbg_frag_t f = specs[i];
	const KbSegX x = bt.segx[i];
	// neighbours: one-base simple pairs on either side (heads have none in front, tails none behind)
	KbSegX sx[3]; int k = 0, at = 0;
	KbSegX dl; dl.s.simple = 1; dl.s.rlen = 1; dl.s.glen = 1; dl.s.rpos = f.rpos - 1; dl.s.gpos = f.gpos - 1; dl.info = KB_SEG_SIMPLE; dl.aux = 0;
	KbSegX dr = dl; dr.s.rpos = f.rpos + f.rlen; dr.s.gpos = f.gpos + f.glen;
	if (f.mode != 1) sx[k++] = dl;
	at = k; sx[k++] = x;
	if (f.mode != 2) sx[k++] = dr;
	kb_dbg_frag_out_t o; memset(&o, 0, sizeof(o));
	KbCigar cg; cg.e = ops + res[i].ops_off; cg.cap = f.rlen + f.glen + 4; cg.n = 0; cg.ovf = false;
	int aln; i64 gf, ge;
	kb_assemble_cigar(pm, bt, sx, k, cg, &aln, &gf, &ge);
	o.info = (int32_t)x.info; o.aux = (int32_t)x.aux; o.score = aln - (k - 1); o.g_first = gf; o.g_end = ge; o.ops_off = res[i].ops_off + (at ? 1u : 0u); o.n_ops = cg.ovf ? -1 : cg.n - (k - 1);
	if (x.info == KB_SEG_JOB) { const KbJob& jb = bt.jobs[x.aux]; o.nruns = jb.nruns; o.ident = jb.ident; o.aligned = jb.aligned; }
	res[i] = o;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
template <class T>
struct DevBuf
{
	T* p = nullptr; size_t n = 0;
	cudaError_t ensure(size_t want) { if (want <= n) return cudaSuccess; if (p) cudaFree(p); p = nullptr; n = 0; cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T)); if (e == cudaSuccess) n = want; return e; }
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Everything one batch in flight owns: a stream, its device arrays and a small pinned block for the counters that come back.
// kb_stage_reads/kb_run/kb_fetch_results use slot 0; kb_map_chunk streams a large chunk through all slots so that the
// H2D copy of one sub-batch, the kernels of the previous one and the D2H copy of the one before overlap.
struct kb_slot
{
	cudaStream_t stream = nullptr; cudaEvent_t ev[10] = {}; cudaEvent_t done = nullptr; cudaEvent_t nw0 = nullptr, nw1 = nullptr;   // nw0..nw1: the nw_alignment solvers alone
	cudaStream_t aux[KB_NW_CLASSES] = {}; cudaEvent_t fork = nullptr, join[KB_NW_CLASSES] = {};   // the nw_alignment size classes run side by side (launch_pipeline)
	KbBatchDev bt; int n_reads = 0; size_t seq_bytes = 0; u64 seq_first = 0; int max_rlen = 0; int first_read = 0;
	DevBuf<u8> seq, scratch, wscratch; DevBuf<u64> seq_off, codes, exc; int packed = 0; u32 n_exc = 0; DevBuf<unsigned long long> work; DevBuf<i32> est, n_hits, n_seeds, n_cands, cand_cap, rescue, slow1, slow2; DevBuf<u32> seed_off, cand_off, cigar, counters, cseg_off, runs; DevBuf<i32> cseg_n; DevBuf<KbSegX> segx; DevBuf<KbJob> jobs; DevBuf<u32> piece_list, part_list; DevBuf<KbPiece> pieces;
	DevBuf<KbHit> hits; DevBuf<KbSeg> segs; DevBuf<KbCand> cands; DevBuf<KbReport> reports; DevBuf<KbReadRes> res; DevBuf<KbPairStat> pstat; DevBuf<kb_aln_t> aln; DevBuf<KbPk> pk; DevBuf<kb_extra_t> extra; DevBuf<KbRTask> rtasks; DevBuf<u32> rjob_first, rjob_count, rslow; size_t cap_rtasks = 0;
	size_t cap_segs = 0, cap_cands = 0, cap_cigar = 0, cap_segx = 0, cap_jobs = 0, cap_pieces = 0, cap_runs = 0, cap_extra = 0, scratch_per_thread = 0; int scratch_threads = 0;
	u32* counters_host = nullptr; unsigned long long* work_dev_host = nullptr;   // pinned: 16 x u32, 8 x u64
	int launches = 0;
	// a chunk in flight through kb_map_chunk_begin / _end: the caller's result buffers and the reads it came from (for a rerun after an arena overflow)
	bool busy = false; kb_results_t* user_out = nullptr; const int32_t* user_est = nullptr; bool user_packed = false; kb_reads_t user_text; kb_reads_packed_t user_pk;
	void release()
	{
		seq.release(); scratch.release(); wscratch.release(); seq_off.release(); codes.release(); exc.release(); work.release(); est.release(); n_hits.release(); n_seeds.release(); n_cands.release(); cand_cap.release();
		rescue.release(); slow1.release(); slow2.release(); seed_off.release(); cand_off.release(); cigar.release(); counters.release(); cseg_off.release(); runs.release();
		cseg_n.release(); segx.release(); jobs.release(); piece_list.release(); part_list.release(); pieces.release(); hits.release(); segs.release(); cands.release(); reports.release(); res.release(); pstat.release(); aln.release(); pk.release(); extra.release(); rtasks.release(); rjob_first.release(); rjob_count.release(); rslow.release();
	}
};

struct kb_ctx
{
	int device = 0; std::string err;
	bool have_index = false; KbIndexDev ix; KbParams pm;
	DevBuf<u32> occ; DevBuf<u64> sa, sa_full, ref64; DevBuf<KbKtab> ktab; DevBuf<u8> pac, lut; DevBuf<i64> chr64; DevBuf<i32> chr32; DevBuf<uint16_t> end_tab;
	int64_t l_pac = 0;
	kb_slot slot[KB_SLOTS];
	DevBuf<u32> chunk_cigar, chunk_cursor;   // cigar arena and cursor shared by the sub-batches of one pipelined chunk
	bool staged = false, ran = false, ran_pipelined = false; u32 n_cigar_last = 0;
	double seg_factor = 32, cigar_factor = 8, scratch_factor = 1, segx_factor = 8, job_factor = 4, run_factor = 96, extra_factor = 0.25, rtask_factor = 0.25;
	float stage_ms[10]; uint64_t work_host[8]; u32 counters_host[KB_NCOUNTERS];
	cudaEvent_t chunk_start = nullptr; int trace = 0;
	int seed_minb = 10;
	int seed_qp = 8, seed_qs = 4, seed_trips = 0;   // trips 0 = 6 (r33 / r35 A/B at C3: 2 trips 3.33 ms, 4 trips 3.27 ms, 6 trips 3.21 ms; 6 on an L2-resident index since r16)   // lane-queue schedule: lanes a pass waits for, lanes a walk waits for, trips per walk (KB_SEED_QP/QS/TRIPS)
	int seg_slab = 256;          // segx slots a warp of k_segments reserves up front (0: none; KB_SEG_SLAB, see kb_alloc_segx; r35 A/B at C3: segments 1.64 -> 1.41 ms)
	int seed_tail = 1;           // a search with at most this many rows left is finished against the text (1: kb_unique_tail only; >1: kb_multi_tail, measured slower at 4..50 in r16, kept as a knob: KB_SEED_TAIL)
	int seed_queue = 1, seed_warps = 148 * 40;   // lane-queue seeding when the full SA is on the device; warps in its grid (KB_SEED_QUEUE, KB_SEED_WARPS)
	bool row32 = false;          // BWT row numbers fit 32 bits: k_fm_seed<.., u32> (set at index upload; KB_ROW64=1 forces the 64-bit kernel)
	int nw_streams = 1;          // 1: the size-class kernels of phase B are forked onto the slot's aux streams and joined before the gather
	int align_warps = KB_ALIGN_WARPS;   // k_nw_warp
	int part_warps = 148 * 40, part_pool = 4096;   // k_align_part: warps in the grid (each with an HBM arena) and shared-memory pool bytes per warp (r14 A/B: 8 warps/SM + 10 KB pool 2.34 ms -> 40 warps/SM + 4 KB 1.34 ms for the align stage at C2)
	int nw_tmax = 0;             // largest side one thread solves (0: KB_NW_TMAX); KB_NW_TMAX=32|64 sends more to the wavefront kernel
	int rf_cand = 256;           // KB_RF_CAND (<= 256)
	int seed_tail_fast = 1;      // kb_unique_tail compares on the read's word boundaries with the strand decided once (KB_SEED_TAIL_FAST=0: the general window loop)
	int seed_stage = 1;          // KB_SEED_STAGE=0: the lane-queue seeding kernel walks a read's packed words in HBM instead of copying them to shared memory first
	int seed_ld_hint = 1;        // Occ blocks, seeding-table and SA entries are loaded with L1::no_allocate in the seeding kernels (r32, C3: seeding 3.47 -> 3.38 ms, L1 hit rate 41 -> 49 %; KB_SEED_LD_HINT=0: plain loads)
	int cand_wide = 1;           // k_cand_heavy_finish: the candidate scan fetches four seeds per round of loads (KB_CAND_WIDE=0: one)
	int fin_local = 1;           // KB_FIN_LOCAL=0: k_finalize works on the reports where they lie
	int rf_reuse = 1, rf_batch = 4;   // k_rescue_fast: the mate's 8-mer index is kept while consecutive windows face the same mate; windows a warp draws per ticket (KB_RF_REUSE, KB_RF_BATCH)
	int rf_stride = 3;           // KB_RF_STRIDE: 3 = k_rescue_fast scans every third window position, 1 = every position
	int part_stack = 24, part_raw = 40;   // KB_PART_STACK / KB_PART_RAW: see KbBatchDev
	int nw_warp_below = 8192;    // a column-tile class (33..64, 65..128) with fewer problems than this is solved by k_nw_warp instead
	int rescue_threads = 64;     // block size of k_rescue_win (32, 64 or 128)
	int cand_heavy = 1;          // items with a long seed list get a warp of their own in k_cand_heavy (KB_CAND_HEAVY=0: everything in k_cand_pair)
	int rescue_fast = 1;         // warp-per-window fast path first (KB_RESCUE_FAST=0: every window through the block-per-window kernel)
	int pipe_min_reads = 262144, pipe_sub_reads = 0;   // chunks of at least pipe_min_reads go through the slot pipeline
	int pipe_first = 0, pipe_grow = 200, pipe_tail = 0;   // sub-batch plan: first size, growth (percent), floor of the halving tail (0: uniform)
	cudaStream_t copy_stream = nullptr;                   // D2H of each sub-batch's cigar range, in retirement order
	int next_slot = 0;                                    // kb_map_chunk_begin: where the search for a free slot starts
};

static int fail(kb_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess)
{
	char buf[512];
	if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e)); else snprintf(buf, sizeof(buf), "%s", what);
	if (c) c->err = buf;
	return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? KB_ENOMEM : KB_ECUDA, #call, e_); } while (0)

extern "C" {

const char* kb_strerror(int code)
{
	switch (code)
	{
	case KB_OK: return "ok";
	case KB_ENODEV: return "no CUDA device (this library has no CPU path)";
	case KB_ECUDA: return "CUDA runtime error";
	case KB_EINVAL: return "invalid argument";
	case KB_ENOMEM: return "out of device memory";
	case KB_ENOINDEX: return "index not uploaded";
	case KB_ECAPACITY: return "result buffer too small";
	case KB_EOVERFLOW: return "device arena overflow persisted after regrowth";
	case KB_ESTATE: return "call out of order";
	default: return "unknown error";
	}
}

const char* kb_last_error(kb_ctx_t* ctx) { return ctx ? ctx->err.c_str() : ""; }

int kb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int kb_init(int device, kb_ctx_t** out)
{
	if (!out) return KB_EINVAL;
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return KB_ENODEV;
	if (device < 0 || device >= n) return KB_EINVAL;
	kb_ctx* ctx = new kb_ctx();
	ctx->device = device;
	memset(&ctx->ix, 0, sizeof(ctx->ix)); memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms)); memset(ctx->work_host, 0, sizeof(ctx->work_host)); memset(ctx->counters_host, 0, sizeof(ctx->counters_host));
	ctx->pm.min_seed = 0; ctx->pm.max_gaps = 5; ctx->pm.max_insert = 1500; ctx->pm.pacbio = 0; ctx->pm.multihit = 0; ctx->pm.paired = 0;
	if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return KB_ECUDA; }
#ifndef KB_EMUL
	{ const char* g = getenv("KB_L2_FETCH"); if (g && (atoi(g) == 32 || atoi(g) == 64 || atoi(g) == 128)) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)); }   // A/B knob: bytes the L2 fetches from HBM per miss
#endif
	for (int k = 0; k < KB_SLOTS; k++)
	{
		kb_slot& sl = ctx->slot[k];
		memset(&sl.bt, 0, sizeof(sl.bt));
		if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess) { kb_destroy(ctx); return KB_ECUDA; }
		for (int i = 0; i < 10; i++) cudaEventCreate(&sl.ev[i]);
		cudaEventCreate(&sl.done); cudaEventCreate(&sl.nw0); cudaEventCreate(&sl.nw1);
		cudaEventCreateWithFlags(&sl.fork, cudaEventDisableTiming);
		for (int i = 0; i < KB_NW_CLASSES; i++) { if (cudaStreamCreateWithFlags(&sl.aux[i], cudaStreamNonBlocking) != cudaSuccess) { kb_destroy(ctx); return KB_ECUDA; } cudaEventCreateWithFlags(&sl.join[i], cudaEventDisableTiming); }
		if (cudaMallocHost((void**)&sl.counters_host, KB_NCOUNTERS * sizeof(u32)) != cudaSuccess || cudaMallocHost((void**)&sl.work_dev_host, 8 * sizeof(unsigned long long)) != cudaSuccess) { kb_destroy(ctx); return KB_ECUDA; }
		memset(sl.counters_host, 0, KB_NCOUNTERS * sizeof(u32)); memset(sl.work_dev_host, 0, 8 * sizeof(unsigned long long));
	}
	cudaEventCreate(&ctx->chunk_start); ctx->trace = getenv("KB_PIPE_TRACE") ? 1 : 0;
	const char* e = getenv("KB_PIPE_MIN_READS"); if (e && atoi(e) > 0) ctx->pipe_min_reads = atoi(e);
	e = getenv("KB_SEED_MINB"); if (e && (atoi(e) == 8 || atoi(e) == 12)) ctx->seed_minb = atoi(e);
	e = getenv("KB_PIPE_SUB_READS"); if (e && atoi(e) > 0) ctx->pipe_sub_reads = atoi(e);
	e = getenv("KB_RESCUE_THREADS"); if (e && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128)) ctx->rescue_threads = atoi(e);
	e = getenv("KB_RESCUE_FAST"); if (e) ctx->rescue_fast = atoi(e) ? 1 : 0;
	e = getenv("KB_CAND_HEAVY"); if (e) ctx->cand_heavy = atoi(e) ? 1 : 0;
	e = getenv("KB_PIPE_FIRST"); if (e && atoi(e) >= -1) ctx->pipe_first = atoi(e);
	e = getenv("KB_PIPE_GROW"); if (e && atoi(e) >= 100) ctx->pipe_grow = atoi(e);
	e = getenv("KB_PIPE_TAIL"); if (e && atoi(e) >= 0) ctx->pipe_tail = atoi(e);
	if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { kb_destroy(ctx); return KB_ECUDA; }
	e = getenv("KB_NW_STREAMS"); if (e) ctx->nw_streams = atoi(e) ? 1 : 0;
	e = getenv("KB_ALIGN_WARPS"); if (e && atoi(e) >= 148 && atoi(e) <= 148 * 64) ctx->align_warps = atoi(e) / 4 * 4;
	e = getenv("KB_NW_TMAX"); if (e && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128)) ctx->nw_tmax = atoi(e);
	e = getenv("KB_SEED_QUEUE"); if (e) ctx->seed_queue = atoi(e) ? 1 : 0;
	e = getenv("KB_SEED_QP"); if (e && atoi(e) >= 1 && atoi(e) <= 32) ctx->seed_qp = atoi(e);
	e = getenv("KB_SEED_QS"); if (e && atoi(e) >= 1 && atoi(e) <= 32) ctx->seed_qs = atoi(e);
	e = getenv("KB_SEED_TRIPS"); if (e && atoi(e) >= 1 && atoi(e) <= 64) ctx->seed_trips = atoi(e);
	e = getenv("KB_SEG_SLAB"); if (e && atoi(e) >= 0 && atoi(e) <= 4096) ctx->seg_slab = atoi(e);
	e = getenv("KB_SEED_TAIL"); if (e && atoi(e) >= 1 && atoi(e) <= 50) ctx->seed_tail = atoi(e);
	e = getenv("KB_SEED_WARPS"); if (e && atoi(e) >= 4 && atoi(e) <= 148 * 64) ctx->seed_warps = atoi(e);
	e = getenv("KB_NW_WARP_BELOW"); if (e && atoi(e) >= 0) ctx->nw_warp_below = atoi(e);
	e = getenv("KB_RF_CAND"); if (e && atoi(e) >= 0 && atoi(e) <= 256) ctx->rf_cand = atoi(e);
	e = getenv("KB_SEED_TAIL_FAST"); if (e) ctx->seed_tail_fast = atoi(e) ? 1 : 0;
	e = getenv("KB_SEED_STAGE"); if (e) ctx->seed_stage = atoi(e) ? 1 : 0;
	e = getenv("KB_SEED_LD_HINT"); if (e) ctx->seed_ld_hint = atoi(e) ? 1 : 0;
	e = getenv("KB_CAND_WIDE"); if (e) ctx->cand_wide = atoi(e) ? 1 : 0;
	e = getenv("KB_FIN_LOCAL"); if (e) ctx->fin_local = atoi(e) ? 1 : 0;
	e = getenv("KB_RF_REUSE"); if (e) ctx->rf_reuse = atoi(e) ? 1 : 0;
	e = getenv("KB_RF_BATCH"); if (e && atoi(e) >= 1 && atoi(e) <= 64) ctx->rf_batch = atoi(e);
	e = getenv("KB_RF_STRIDE"); if (e && (atoi(e) == 1 || atoi(e) == 3)) ctx->rf_stride = atoi(e);
	e = getenv("KB_PART_STACK"); if (e && atoi(e) >= 1) ctx->part_stack = atoi(e);
	e = getenv("KB_PART_RAW"); if (e && atoi(e) >= 1) ctx->part_raw = atoi(e);
	e = getenv("KB_PART_WARPS"); if (e && atoi(e) >= 148 && atoi(e) <= 148 * 64) ctx->part_warps = atoi(e) / 4 * 4;
	e = getenv("KB_PART_POOL"); if (e && atoi(e) >= 1024 && atoi(e) <= 11264) ctx->part_pool = atoi(e) / 16 * 16;
	*out = ctx;
	return KB_OK;
}

void kb_destroy(kb_ctx_t* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	for (int k = 0; k < KB_SLOTS; k++) if (ctx->slot[k].stream) cudaStreamSynchronize(ctx->slot[k].stream);
	if (ctx->chunk_start) cudaEventDestroy(ctx->chunk_start);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	ctx->occ.release(); ctx->ktab.release(); ctx->ref64.release(); ctx->sa.release(); ctx->sa_full.release(); ctx->pac.release(); ctx->lut.release(); ctx->chr64.release(); ctx->chr32.release(); ctx->end_tab.release();
	ctx->chunk_cigar.release(); ctx->chunk_cursor.release();
	for (int k = 0; k < KB_SLOTS; k++)   // every handle may be null: kb_init calls this on its failure paths
	{
		kb_slot& sl = ctx->slot[k];
		sl.release();
		for (int i = 0; i < 10; i++) if (sl.ev[i]) cudaEventDestroy(sl.ev[i]);
		if (sl.done) cudaEventDestroy(sl.done);
		if (sl.nw0) cudaEventDestroy(sl.nw0);
		if (sl.nw1) cudaEventDestroy(sl.nw1);
		if (sl.fork) cudaEventDestroy(sl.fork);
		for (int i = 0; i < KB_NW_CLASSES; i++) { if (sl.join[i]) cudaEventDestroy(sl.join[i]); if (sl.aux[i]) cudaStreamDestroy(sl.aux[i]); }
		if (sl.counters_host) cudaFreeHost(sl.counters_host);
		if (sl.work_dev_host) cudaFreeHost(sl.work_dev_host);
		if (sl.stream) cudaStreamDestroy(sl.stream);
	}
	delete ctx;
}

static int derive_min_seed(int64_t l_pac)   // src/Mapping.cpp:645
{
	int m; double two_g = (double)(l_pac * 2);
	for (m = 13; m < 16; m++) if (two_g < pow(4, m)) break;
	return m;
}

int kb_upload_index(kb_ctx_t* ctx, const kb_index_host_t* h, int expand_sa)
{
	if (!ctx || !h || !h->bwt || !h->sa || !h->pac || h->n_chr <= 0 || !h->chr_len || h->sa_intv <= 0 || (h->sa_intv & (h->sa_intv - 1))) return fail(ctx, KB_EINVAL, "kb_upload_index: bad index description");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t stream = ctx->slot[0].stream;
	for (int c = 1; c <= 4; c++) if (h->L2[c] - h->L2[c - 1] >= 0xFFFFFFFFull) return fail(ctx, KB_EINVAL, "kb_upload_index: a base count exceeds 2^32-1 (u32 Occ layout)");
	if (h->seq_len + 2 >= (1ull << 40)) return fail(ctx, KB_EINVAL, "kb_upload_index: text longer than 2^40 (row numbers are packed into 40 bits)");
	KbIndexDev& ix = ctx->ix; memset(&ix, 0, sizeof(ix));
	ix.primary = h->primary; for (int i = 0; i < 5; i++) ix.L2[i] = h->L2[i]; ix.seq_len = h->seq_len;
	// Occ re-blocking on the device
	u64 n_new = (h->seq_len >> 6) + 2;
	{
		DevBuf<u32> raw;
		CK(raw.ensure(h->bwt_words)); CK(ctx->occ.ensure(n_new * 8));
		CK(cudaMemcpyAsync(raw.p, h->bwt, h->bwt_words * 4, cudaMemcpyHostToDevice, stream));
		KB_LAUNCH(k_reblock, (unsigned)((n_new + 255) / 256), 256, stream, raw.p, h->bwt_words, n_new, ctx->occ.p);
		CK(cudaGetLastError()); CK(cudaStreamSynchronize(stream));
		raw.release();
	}
	ix.occ = ctx->occ.p; ix.n_blocks = n_new;
	CK(ctx->sa.ensure(h->n_sa)); CK(cudaMemcpyAsync(ctx->sa.p, h->sa, h->n_sa * 8, cudaMemcpyHostToDevice, stream));
	ix.sa = ctx->sa.p; ix.n_sa = h->n_sa; ix.sa_intv = h->sa_intv; ix.sa_full = nullptr;
	size_t pac_bytes = (size_t)(h->l_pac / 4 + 1);
	CK(ctx->pac.ensure(pac_bytes)); CK(cudaMemcpyAsync(ctx->pac.p, h->pac, pac_bytes, cudaMemcpyHostToDevice, stream));
	ix.pac = ctx->pac.p; ix.G = h->l_pac; ix.G2 = h->l_pac * 2; ctx->l_pac = h->l_pac;
	{
		u64 words = (pac_bytes + 7) / 8 + 2;
		CK(ctx->ref64.ensure(words));
		KB_LAUNCH(k_ref64, (unsigned)((words + 255) / 256), 256, stream, ctx->pac.p, (u64)pac_bytes, words, ctx->ref64.p);
		CK(cudaGetLastError());
		ix.ref64 = ctx->ref64.p;
	}
	// chromosome tables: ChrLocMap (src/bwt_index.cpp:250-251) as a sorted key array
	int nc = h->n_chr, ne = 2 * nc;
	std::vector<i64> t64((size_t)ne + 3 * nc); std::vector<i32> t32(ne);
	i64* key = t64.data(); i64* fwd = key + ne; i64* rev = fwd + nc; i64* len = rev + nc;
	i64 total = 0;
	for (int i = 0; i < nc; i++) { len[i] = h->chr_len[i]; fwd[i] = total; total += len[i]; rev[i] = ix.G2 - total; }
	for (int i = 0; i < nc; i++) { key[i] = fwd[i] + len[i] - 1; t32[i] = i; }                       // forward ends ascend with i
	for (int i = 0; i < nc; i++) { int c = nc - 1 - i; key[nc + i] = rev[c] + len[c] - 1; t32[nc + i] = c; }   // reverse ends ascend with descending i
	CK(ctx->chr64.ensure(t64.size())); CK(ctx->chr32.ensure(t32.size()));
	CK(cudaMemcpyAsync(ctx->chr64.p, t64.data(), t64.size() * 8, cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(ctx->chr32.p, t32.data(), t32.size() * 4, cudaMemcpyHostToDevice, stream));
	ix.end_tab = nullptr; ix.end_shift = 0;
	if (ne <= 65535 && !(getenv("KB_CHR_TAB") && !atoi(getenv("KB_CHR_TAB"))))   // bucket table for kb_chr_lookup
	{
		int sh = 10; while (((u64)ix.G2 >> sh) + 2 > 8192) sh++;
		const size_t nt = (size_t)((u64)ix.G2 >> sh) + 2;
		std::vector<uint16_t> tab(nt);
		int at = 0;
		for (size_t b = 0; b < nt; b++) { const i64 start = (i64)b << sh; while (at < ne && key[at] < start) at++; tab[b] = (uint16_t)at; }
		CK(ctx->end_tab.ensure(nt));
		CK(cudaMemcpyAsync(ctx->end_tab.p, tab.data(), nt * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));
		CK(cudaStreamSynchronize(stream));   // tab is a local
		ix.end_tab = ctx->end_tab.p; ix.end_shift = sh;
	}
	ix.n_chr = nc; ix.n_ends = ne; ix.end_key = ctx->chr64.p; ix.chr_fwd = ctx->chr64.p + ne; ix.chr_rev = ix.chr_fwd + nc; ix.chr_len = ix.chr_rev + nc; ix.end_chr = ctx->chr32.p;
	// MAPQ table with the reference expression (src/Mapping.cpp:172), evaluated by the host libm exactly like the reference
	const int lut_scores = 1 << 16;
	std::vector<u8> lut((size_t)lut_scores * 5, 0);
	for (int s = 1; s < lut_scores; s++)
		for (int d = 1; d <= 5 && d < s; d++)
		{
			int score = s, sub = s - d;
			int q = (int)(30 * (1 - (float)(score - sub) / score) * log(score) + 0.4999);
			lut[(size_t)s * 5 + d - 1] = (u8)(q > 60 ? 60 : (q < 0 ? 0 : q));
		}
	CK(ctx->lut.ensure(lut.size())); CK(cudaMemcpyAsync(ctx->lut.p, lut.data(), lut.size(), cudaMemcpyHostToDevice, stream));
	ix.mapq_lut = ctx->lut.p; ix.mapq_lut_scores = lut_scores;
	CK(cudaStreamSynchronize(stream));
	{
		// seeding table: the smallest K with 4^K >= 2G, at most 14 (4.3 GB of the 180 GB): past it intervals are narrow, so
		// nearly every remaining extension step is the one-block case of kb_extend. Never longer than MinSeedLength.
		int K = 1; while (K < 14 && (1ull << (2 * K)) < h->seq_len) K++;
		const int ms = derive_min_seed(h->l_pac); if (K > ms) K = ms;
		{ const char* e = getenv("KB_KTAB_K"); if (e && atoi(e) >= 0 && atoi(e) <= 16) K = atoi(e) < ms ? atoi(e) : ms; }   // 16 bytes x 4^K: 4.3 GB at 14, 17 GB at 15, 69 GB at 16
		if (K >= 4)
		{
			u64 ne2 = 1ull << (2 * K);
			CK(ctx->ktab.ensure(ne2));
			ix.ktab = nullptr; ix.ktab_k = 0;
			KB_LAUNCH(k_build_ktab, (unsigned)((ne2 + 255) / 256), 256, stream, ix, K, ctx->ktab.p);
			CK(cudaGetLastError()); CK(cudaStreamSynchronize(stream));
			ix.ktab = ctx->ktab.p; ix.ktab_k = K;
		}
	}
	if (expand_sa == 2)   // auto: only when the full SA leaves room for the batches
	{
		size_t free_b = 0, total_b = 0;
#ifndef KB_EMUL
		if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
#endif
		expand_sa = ((h->seq_len + 1) * 8ull + (48ull << 30) <= (unsigned long long)free_b) ? 1 : 0;
	}
	if (expand_sa)
	{
		CK(ctx->sa_full.ensure(h->seq_len + 1));
		KB_LAUNCH(k_expand_sa, (unsigned)((h->seq_len + 256) / 256), 256, stream, ix, ctx->sa_full.p);
		CK(cudaGetLastError()); CK(cudaStreamSynchronize(stream));
		ix.sa_full = ctx->sa_full.p;
	}
	if (ctx->pm.min_seed <= 0) ctx->pm.min_seed = derive_min_seed(h->l_pac);
	ctx->row32 = (h->seq_len + 2 < 0xFFFFFFFFull) && !(getenv("KB_ROW64") && atoi(getenv("KB_ROW64")));
	ctx->have_index = true;
	return KB_OK;
}

// A second device gets the index from the first one's HBM instead of over PCIe from the host: everything kb_upload_index built
// (re-blocked Occ, seeding table, full SA, packed reference, tables) is copied device to device -- over NVLink when the two
// devices are peers -- which also skips the table build and the SA expansion.
int kb_clone_index(kb_ctx_t* dst, kb_ctx_t* src)
{
	kb_ctx* ctx = dst;
	if (!dst || !src || dst == src) return fail(ctx, KB_EINVAL, "kb_clone_index: bad arguments");
	if (!src->have_index) return fail(ctx, KB_ENOINDEX, "kb_clone_index: the source context has no index");
	CK(cudaSetDevice(dst->device));
#ifndef KB_EMUL
	if (dst->device != src->device)
	{
		int can = 0; cudaDeviceCanAccessPeer(&can, dst->device, src->device);
		if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0); if (e != cudaSuccess) cudaGetLastError(); }   // already enabled is fine
	}
#endif
	cudaStream_t stream = dst->slot[0].stream;
	auto copy = [&](auto& d, const auto& sbuf) -> int {
		if (sbuf.n == 0) return KB_OK;
		CK(d.ensure(sbuf.n));
#ifndef KB_EMUL
		CK(cudaMemcpyPeerAsync(d.p, dst->device, sbuf.p, src->device, sbuf.n * sizeof(*sbuf.p), stream));
#else
		memcpy(d.p, sbuf.p, sbuf.n * sizeof(*sbuf.p));
#endif
		return KB_OK;
	};
	int rc;
	if ((rc = copy(dst->occ, src->occ)) || (rc = copy(dst->sa, src->sa)) || (rc = copy(dst->sa_full, src->sa_full)) || (rc = copy(dst->ktab, src->ktab)) || (rc = copy(dst->ref64, src->ref64))
	    || (rc = copy(dst->pac, src->pac)) || (rc = copy(dst->lut, src->lut)) || (rc = copy(dst->chr64, src->chr64)) || (rc = copy(dst->chr32, src->chr32)) || (rc = copy(dst->end_tab, src->end_tab))) return rc;
	CK(cudaStreamSynchronize(stream));
	KbIndexDev ix = src->ix;
	ix.occ = dst->occ.p; ix.sa = dst->sa.p; ix.sa_full = src->ix.sa_full ? dst->sa_full.p : nullptr; ix.ktab = src->ix.ktab ? dst->ktab.p : nullptr;
	ix.pac = dst->pac.p; ix.ref64 = dst->ref64.p; ix.mapq_lut = dst->lut.p; ix.end_tab = src->ix.end_tab ? dst->end_tab.p : nullptr;
	const int nc = ix.n_chr, ne = ix.n_ends;
	ix.end_key = dst->chr64.p; ix.chr_fwd = dst->chr64.p + ne; ix.chr_rev = ix.chr_fwd + nc; ix.chr_len = ix.chr_rev + nc; ix.end_chr = dst->chr32.p;
	dst->ix = ix; dst->l_pac = src->l_pac; dst->row32 = src->row32; dst->have_index = true;
	if (dst->pm.min_seed <= 0) dst->pm.min_seed = derive_min_seed(dst->l_pac);
	return KB_OK;
}

int kb_set_params(kb_ctx_t* ctx, const kb_params_t* p)
{
	if (!ctx || !p) return KB_EINVAL;
	ctx->pm.max_gaps = p->max_gaps < 0 ? 0 : p->max_gaps; ctx->pm.max_insert = p->max_insert > 0 ? p->max_insert : 1500;
	ctx->pm.pacbio = p->pacbio ? 1 : 0; ctx->pm.multihit = p->multihit ? 1 : 0; ctx->pm.paired = (p->paired && !p->pacbio) ? 1 : 0;
	ctx->pm.min_seed = p->min_seed_len > 0 ? p->min_seed_len : (ctx->have_index ? derive_min_seed(ctx->l_pac) : 0);
	return KB_OK;
}

int kb_get_min_seed_len(kb_ctx_t* ctx) { return ctx ? ctx->pm.min_seed : KB_EINVAL; }
void* kb_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
#ifndef KB_EMUL
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
#else
	p = malloc(bytes ? bytes : 1);
#endif
	return p;
}
void kb_host_free(void* p)
{
	if (!p) return;
#ifndef KB_EMUL
	cudaFreeHost(p);
#else
	free(p);
#endif
}
int kb_host_register(void* p, uint64_t bytes)
{
	if (!p || !bytes) return KB_EINVAL;
#ifndef KB_EMUL
	if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return KB_ECUDA; }
#endif
	return KB_OK;
}
void kb_host_unregister(void* p)
{
#ifndef KB_EMUL
	if (p) { if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError(); }
#else
	(void)p;
#endif
}
void* kb_cuda_stream(kb_ctx_t* ctx) { return ctx ? (void*)ctx->slot[0].stream : nullptr; }

// device arrays of one slot for its n_reads / max_rlen; shared != 0: cigar elements go to the chunk-wide arena
static int alloc_batch(kb_ctx* ctx, kb_slot& sl, int shared)
{
	size_t n = (size_t)sl.n_reads; KbBatchDev& bt = sl.bt;
	int L = sl.max_rlen > 0 ? sl.max_rlen : 1;
	int max_hits = ctx->pm.pacbio ? L / ctx->pm.min_seed + 2 : L / (ctx->pm.min_seed + 1) + 2;
	sl.cap_segs = (size_t)(ctx->seg_factor * (double)n) + (ctx->pm.pacbio ? (size_t)n * (size_t)(L / 8 + 64) : 0) + 65536;
	if (sl.cap_segs > 0xF0000000ull) sl.cap_segs = 0xF0000000ull;
	sl.cap_cands = 2 * sl.cap_segs + 2 * n + 1024;
	if (sl.cap_cands > 0xF0000000ull) sl.cap_cands = 0xF0000000ull;
	sl.cap_cigar = (size_t)(ctx->cigar_factor * (double)n) + (ctx->pm.pacbio ? (size_t)n * (size_t)(L / 2) : 0) + 65536;
	sl.cap_segx = (size_t)(ctx->segx_factor * (double)n) + (ctx->pm.pacbio ? (size_t)n * (size_t)(L / 16 + 64) : 0) + 65536 + (size_t)ctx->seg_slab * (n / 32 + 4);
	sl.cap_jobs = (size_t)(ctx->job_factor * (double)n) + (ctx->pm.pacbio ? (size_t)n * (size_t)(L / 32 + 32) : 0) + 65536;
	sl.cap_pieces = 2 * sl.cap_jobs;
	sl.cap_runs = (size_t)(ctx->run_factor * (double)n) + (ctx->pm.pacbio ? (size_t)n * (size_t)(3 * L) : 0) + (1 << 20);
	if (sl.cap_runs > 0xF0000000ull) sl.cap_runs = 0xF0000000ull;
	// per-thread scratch of the arena kernels: NW traceback (2 bit / cell) dominates
	double side = L < 3100 ? L + 64 : 3100 + 64;
	size_t per = (size_t)(side * side / 4) + (size_t)L * 160 + (64 << 10);
	if (ctx->pm.pacbio) per += (size_t)L * 64;
	per = (size_t)((double)per * ctx->scratch_factor); per = (per + 255) & ~(size_t)255;
	size_t budget = (size_t)24 << 30;
	size_t threads = budget / per; if (threads > 148 * 1024) threads = 148 * 1024;
	if (threads > n) threads = n;
	threads = (threads + KB_BLOCK - 1) / KB_BLOCK * KB_BLOCK; if (threads < KB_BLOCK) threads = KB_BLOCK;
	sl.scratch_per_thread = per; sl.scratch_threads = (int)threads;
	// the warp-per-job kernels get their own arenas (one worst-case problem each), independent of the number of reads
	const int wwarps = ctx->align_warps > ctx->part_warps ? ctx->align_warps : ctx->part_warps;
	CK(sl.hits.ensure(n * max_hits)); CK(sl.n_hits.ensure(n)); CK(sl.n_seeds.ensure(n)); CK(sl.seed_off.ensure(n));
	CK(sl.segs.ensure(sl.cap_segs)); CK(sl.cands.ensure(sl.cap_cands)); CK(sl.reports.ensure(sl.cap_cands));
	CK(sl.n_cands.ensure(n)); CK(sl.cand_off.ensure(n)); CK(sl.cand_cap.ensure(n)); CK(sl.rescue.ensure(n / 2 + 1));
	sl.cap_rtasks = ctx->pm.paired ? (size_t)(ctx->rtask_factor * (double)n) + 65536 : 1;
	CK(sl.rtasks.ensure(sl.cap_rtasks)); CK(sl.rslow.ensure(sl.cap_rtasks)); CK(sl.rjob_first.ensure(n / 2 + 1)); CK(sl.rjob_count.ensure(n / 2 + 1));
	bt.segx_slab = nullptr; bt.rtasks = sl.rtasks.p; bt.cap_rtasks = (u32)sl.cap_rtasks; bt.rjob_first = sl.rjob_first.p; bt.rjob_count = sl.rjob_count.p; bt.rslow = sl.rslow.p;
	CK(sl.res.ensure(n)); CK(sl.pstat.ensure(n / 2 + 1)); CK(sl.aln.ensure(n));
	if (!shared) CK(sl.cigar.ensure(sl.cap_cigar));
	sl.cap_extra = ctx->pm.multihit ? (size_t)(ctx->extra_factor * (double)n) + 65536 : 0;
	if (sl.cap_extra > 0xF0000000ull) sl.cap_extra = 0xF0000000ull;
	if (sl.cap_extra) CK(sl.extra.ensure(sl.cap_extra));
	bt.extra = sl.extra.p; bt.cap_extra = (u32)sl.cap_extra;
	CK(sl.segx.ensure(sl.cap_segx)); CK(sl.jobs.ensure(sl.cap_jobs)); CK(sl.pieces.ensure(sl.cap_pieces)); CK(sl.piece_list.ensure(sl.cap_pieces * KB_NW_CLASSES)); CK(sl.part_list.ensure(sl.cap_jobs)); CK(sl.runs.ensure(sl.cap_runs)); CK(sl.cseg_off.ensure(sl.cap_cands)); CK(sl.cseg_n.ensure(sl.cap_cands));
	CK(sl.counters.ensure(KB_NCOUNTERS)); CK(sl.work.ensure(8)); CK(sl.scratch.ensure(per * threads)); CK(sl.wscratch.ensure(per * (size_t)wwarps));
	CK(sl.pk.ensure((sl.seq_bytes >> 5) + n + 4)); CK(sl.slow1.ensure(n + 1)); CK(sl.slow2.ensure(n + 1));
	// reads keep their chunk-wide offsets: the device copies start at seq_first, so the base pointers are shifted back by it
	bt.n_reads = sl.n_reads; bt.seq = sl.seq.p - sl.seq_first; bt.seq_off = sl.seq_off.p; bt.est = sl.est.p; bt.pk = sl.pk.p - (sl.seq_first >> 5); bt.pk_wpr = (L + 31) / 32;
	bt.hits = sl.hits.p; bt.max_hits = max_hits; bt.n_hits = sl.n_hits.p; bt.n_seeds = sl.n_seeds.p; bt.seed_off = sl.seed_off.p;
	bt.segs = sl.segs.p; bt.cap_segs = (u32)sl.cap_segs; bt.cands = sl.cands.p; bt.cap_cands = (u32)sl.cap_cands; bt.n_cands = sl.n_cands.p;
	bt.cand_off = sl.cand_off.p; bt.cand_cap = sl.cand_cap.p; bt.rescue_list = sl.rescue.p; bt.slow_list = sl.slow1.p; bt.slow_list2 = sl.slow2.p; bt.reports = sl.reports.p; bt.res = sl.res.p; bt.pstat = sl.pstat.p;
	bt.segx = sl.segx.p; bt.cap_segx = (u32)sl.cap_segx; bt.cseg_off = sl.cseg_off.p; bt.cseg_n = sl.cseg_n.p; bt.jobs = sl.jobs.p; bt.cap_jobs = (u32)sl.cap_jobs; bt.pieces = sl.pieces.p; bt.cap_pieces = (u32)sl.cap_pieces; bt.piece_list = sl.piece_list.p; bt.part_list = sl.part_list.p; bt.runs = sl.runs.p; bt.cap_runs = (u32)sl.cap_runs;
	if (shared) { bt.cigar = ctx->chunk_cigar.p; bt.cap_cigar = (u32)ctx->chunk_cigar.n; bt.cig_cursor = ctx->chunk_cursor.p; }
	else { bt.cigar = sl.cigar.p; bt.cap_cigar = (u32)sl.cap_cigar; bt.cig_cursor = sl.counters.p + 2; }
	bt.scratch = sl.scratch.p; bt.scratch_per_thread = per; bt.scratch_threads = (int)threads;
	bt.wscratch = sl.wscratch.p; bt.wscratch_per_warp = per; bt.wscratch_warps = wwarps;
	bt.max_rlen = L; bt.nw_tmax = ctx->nw_tmax > 0 ? ctx->nw_tmax : KB_NW_TMAX; bt.nw_warp_below = ctx->nw_warp_below; bt.rf_cand = ctx->rf_cand; bt.rf_stride = ctx->rf_stride; bt.fin_local = ctx->fin_local; bt.rf_reuse = ctx->rf_reuse; bt.rf_batch = ctx->rf_batch; bt.part_stack = ctx->part_stack; bt.part_raw = ctx->part_raw; bt.seg_cap = 0; bt.kmer_cap = 0; bt.counters = sl.counters.p; bt.work = sl.work.p;
	return KB_OK;
}

// the reads of a chunk as either entry point hands them over
struct KbReadSrc { int n; const u8* seq; const u64* seq_off; const kb_reads_packed_t* pk; };
static KbReadSrc read_src(const kb_reads_t* in) { KbReadSrc s; s.n = in->n_reads; s.seq = in->seq; s.seq_off = in->seq_off; s.pk = nullptr; return s; }
static KbReadSrc read_src_packed(const kb_reads_packed_t* in) { KbReadSrc s; s.n = in->n_reads; s.seq = nullptr; s.seq_off = in->seq_off; s.pk = in; return s; }

// H2D of reads [first, first + count) of the chunk into a slot (asynchronous on the slot's stream): the characters, or the packed
// words and the exception entries of that range (k_unpack rebuilds the characters on the device)
static int stage_slot(kb_ctx* ctx, kb_slot& sl, const KbReadSrc& in, int first, int count, const int32_t* est)
{
	size_t n = (size_t)count;
	sl.n_reads = count; sl.first_read = first;
	sl.seq_first = n ? in.seq_off[first] : 0;
	sl.seq_bytes = n ? (size_t)(in.seq_off[first + n] - sl.seq_first) : 0;
	int L = 0; for (size_t i = 0; i < n; i++) { u64 l = in.seq_off[first + i + 1] - in.seq_off[first + i]; if (l > (in.pk ? 0xFFFFFFull : 0x7FFFFFF0ull)) return fail(ctx, KB_EINVAL, "read too long"); if ((int)l > L) L = (int)l; }
	sl.max_rlen = L; sl.packed = in.pk ? 1 : 0; sl.n_exc = 0;
	CK(sl.seq.ensure(sl.seq_bytes + 64)); CK(sl.seq_off.ensure(n + 1)); CK(sl.est.ensure(n / 2 + 1));
	if (n)
	{
		if (!in.pk) CK(cudaMemcpyAsync(sl.seq.p, in.seq + sl.seq_first, sl.seq_bytes, cudaMemcpyHostToDevice, sl.stream));
		else
		{
			const u64 w0 = (sl.seq_first >> 5) + (u64)first, nw = (in.seq_off[first + n] >> 5) - (sl.seq_first >> 5) + n + 1;
			if (w0 + nw > in.pk->n_words) return fail(ctx, KB_EINVAL, "packed reads: n_words is smaller than the layout needs");
			CK(sl.codes.ensure(nw));
			CK(cudaMemcpyAsync(sl.codes.p, in.pk->code + w0, nw * 8, cudaMemcpyHostToDevice, sl.stream));
			const u64* e0 = std::lower_bound(in.pk->exc, in.pk->exc + in.pk->n_exc, (u64)first << 32);
			const u64* e1 = std::lower_bound(e0, in.pk->exc + in.pk->n_exc, (u64)(first + n) << 32);
			sl.n_exc = (u32)(e1 - e0);
			if (sl.n_exc) { CK(sl.exc.ensure(sl.n_exc)); CK(cudaMemcpyAsync(sl.exc.p, e0, (size_t)sl.n_exc * 8, cudaMemcpyHostToDevice, sl.stream)); }
		}
		CK(cudaMemcpyAsync(sl.seq_off.p, in.seq_off + first, (n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
		if (ctx->pm.paired) CK(cudaMemcpyAsync(sl.est.p, est + first / 2, (n / 2) * 4, cudaMemcpyHostToDevice, sl.stream));
	}
	return KB_OK;
}

static int stage_reads(kb_ctx* ctx, const KbReadSrc& in, const int32_t* est, const char* who)
{
	if (in.n < 0 || (in.n > 0 && ((!in.seq && !in.pk) || !in.seq_off || (in.pk && (!in.pk->code || (in.pk->n_exc && !in.pk->exc)))))) return fail(ctx, KB_EINVAL, who);
	if (!ctx->have_index) return fail(ctx, KB_ENOINDEX, "kb_stage_reads: no index");
	if (ctx->pm.paired && ((in.n & 1) || (in.n > 0 && !est))) return fail(ctx, KB_EINVAL, "kb_stage_reads: paired chunks need an even read count and one EstDistance per pair");
	for (int k = 0; k < KB_SLOTS; k++) if (ctx->slot[k].busy) return fail(ctx, KB_ESTATE, "a chunk is in flight (kb_map_chunk_end it first)");
	CK(cudaSetDevice(ctx->device));
	ctx->staged = false; ctx->ran = false; ctx->ran_pipelined = false;
	int rc = stage_slot(ctx, ctx->slot[0], in, 0, in.n, est); if (rc) return rc;
	ctx->staged = true;
	return KB_OK;
}

int kb_stage_reads(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est)
{
	if (!ctx || !in) return fail(ctx, KB_EINVAL, "kb_stage_reads: bad arguments");
	return stage_reads(ctx, read_src(in), est, "kb_stage_reads: bad arguments");
}
int kb_stage_reads_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est)
{
	if (!ctx || !in) return fail(ctx, KB_EINVAL, "kb_stage_reads_packed: bad arguments");
	return stage_reads(ctx, read_src_packed(in), est, "kb_stage_reads_packed: bad arguments");
}

// the slot's reads in both device forms: characters -> packed words (k_pack), or packed words (+ exceptions) -> characters (k_unpack)
static void launch_pack(kb_slot& sl)
{
	KbBatchDev& bt = sl.bt; cudaStream_t s = sl.stream; const int n = sl.n_reads;
	const unsigned g = (unsigned)(((long long)n * bt.pk_wpr + KB_BLOCK - 1) / KB_BLOCK);
	if (sl.packed)
	{
		KB_LAUNCH(k_unpack, g, KB_BLOCK, s, bt, sl.codes.p, sl.seq.p); sl.launches++;
		if (sl.n_exc) { KB_LAUNCH(k_unpack_exc, (unsigned)((sl.n_exc + KB_BLOCK - 1) / KB_BLOCK), KB_BLOCK, s, bt, sl.exc.p, sl.n_exc, (u32)sl.first_read, sl.seq.p); sl.launches++; }
	}
	else { KB_LAUNCH(k_pack, g, KB_BLOCK, s, bt); sl.launches++; }
}

// phase B of a slot's batch: 8-mer partition, the nw_alignment size classes, gather (kb_align.cuh "phase B")
static int launch_phase_b(kb_ctx* ctx, kb_slot& sl)
{
	KbBatchDev& bt = sl.bt; const KbIndexDev& ix = ctx->ix; const KbParams& pm = ctx->pm; cudaStream_t s = sl.stream;
	unsigned g_warp = (unsigned)(ctx->align_warps * 32 / KB_BLOCK);
#ifndef KB_EMUL
	k_align_part<<<(unsigned)(ctx->part_warps * 32 / KB_BLOCK), KB_BLOCK, (size_t)(KB_BLOCK / 32) * (size_t)ctx->part_pool, s>>>(ix, pm, bt, ctx->part_pool);
#else
	KB_LAUNCH(k_align_part, g_warp, KB_BLOCK, s, ix, pm, bt, ctx->part_pool);
#endif
	sl.launches++;
	CK(cudaEventRecord(sl.nw0, s));
	// The size classes are independent and each is latency-bound on its own (one problem per thread: a launch lasts as long as
	// its longest chain of cells), so they run side by side on the slot's aux streams, heaviest first, and join before the gather.
	{
		cudaStream_t q[KB_NW_CLASSES];
		for (int i = 0; i < KB_NW_CLASSES; i++) q[i] = ctx->nw_streams ? sl.aux[i] : s;
		if (ctx->nw_streams) { CK(cudaEventRecord(sl.fork, s)); for (int i = 0; i < KB_NW_CLASSES; i++) CK(cudaStreamWaitEvent(q[i], sl.fork, 0)); }
		KB_LAUNCH((k_nw_tile<5, 32, 128, 4>), 148 * 4, KB_BLOCK, q[5], ix, pm, bt); sl.launches++;
		KB_LAUNCH((k_nw_tile<4, 32, 64, 2>), 148 * 4, KB_BLOCK, q[4], ix, pm, bt); sl.launches++;
		KB_LAUNCH(k_nw_warp, g_warp, KB_BLOCK, q[6], ix, pm, bt); sl.launches++;
		KB_LAUNCH((k_nw_tile<3, 32, 32, 1>), 148 * 8, KB_BLOCK, q[3], ix, pm, bt); sl.launches++;
		KB_LAUNCH((k_nw_tile<2, 24, 32, 1>), 148 * 8, KB_BLOCK, q[2], ix, pm, bt); sl.launches++;
		KB_LAUNCH((k_nw_tile<1, 16, 32, 1>), 148 * 8, KB_BLOCK, q[1], ix, pm, bt); sl.launches++;
		KB_LAUNCH((k_nw_tile<0, 8, 32, 1>), 148 * 8, KB_BLOCK, q[0], ix, pm, bt); sl.launches++;
		if (ctx->nw_streams) for (int i = 0; i < KB_NW_CLASSES; i++) { CK(cudaEventRecord(sl.join[i], q[i])); CK(cudaStreamWaitEvent(s, sl.join[i], 0)); }
	}
	CK(cudaEventRecord(sl.nw1, s));
	KB_LAUNCH(k_align_gather, 148 * 4, KB_BLOCK, s, bt); sl.launches++;
	return KB_OK;
}

static int launch_pipeline(kb_ctx* ctx, kb_slot& sl)
{
	KbBatchDev& bt = sl.bt; const KbIndexDev& ix = ctx->ix; const KbParams& pm = ctx->pm;
	int n = sl.n_reads; cudaStream_t s = sl.stream;
	CK(cudaMemsetAsync(sl.counters.p, 0, KB_NCOUNTERS * sizeof(u32), s)); CK(cudaMemsetAsync(sl.work.p, 0, 8 * sizeof(u64), s));
	unsigned g_reads = (unsigned)((n + KB_BLOCK - 1) / KB_BLOCK);
	unsigned g_items = pm.paired ? (unsigned)((n / 2 + KB_BLOCK - 1) / KB_BLOCK) : g_reads;
	unsigned g_hits = (unsigned)(((long long)n * bt.max_hits + KB_BLOCK - 1) / KB_BLOCK);
	unsigned g_slow = g_reads < 148u * 16u ? g_reads : 148u * 16u;   // arena kernels: one thread per read up to a full machine, slices cut on the device
	sl.launches = 0;
	CK(cudaEventRecord(sl.ev[0], s));
	launch_pack(sl);
	KbIndexDev ixs = ix; ixs.ld_hint = ctx->seed_ld_hint | (ctx->seed_tail_fast ? 0 : 2);   // the seeding kernels' loads of Occ blocks, table and SA entries (kb_load_blk)
	if (ctx->seed_queue && ix.sa_full != nullptr)
	{
		// with the full SA most searches finish against the text and reads differ widely in work: lane queue (kb_seed_lane)
		// two reads per lane at least, so that the queue has something to balance -- but a long read is thousands of dependent steps on its
		// own: those get a lane each, as many warps resident as there are reads (C5: 50 k reads were 782 warps = 5 per SM)
		unsigned warps = sl.max_rlen >= 1000 ? (unsigned)((n + 31) / 32) : (unsigned)((n + 63) / 64);
		if (warps > (unsigned)ctx->seed_warps) warps = (unsigned)ctx->seed_warps; if (warps < 4) warps = 4;
		const unsigned gq = (warps + 3) / 4;
		const size_t seed_smem = ctx->seed_stage ? (size_t)KB_BLOCK * KB_SEED_STAGE * sizeof(KbPk) : 0;
		if (ctx->row32) { if (ctx->seed_minb == 12) { KB_LAUNCH_SMEM((k_fm_seed_q<12, u32>), gq, KB_BLOCK, seed_smem, s, ixs, pm, bt, ctx->seed_qp, ctx->seed_qs, ctx->seed_trips > 0 ? ctx->seed_trips : 6, ctx->seed_tail, ctx->seed_stage); } else { KB_LAUNCH_SMEM((k_fm_seed_q<10, u32>), gq, KB_BLOCK, seed_smem, s, ixs, pm, bt, ctx->seed_qp, ctx->seed_qs, ctx->seed_trips > 0 ? ctx->seed_trips : 6, ctx->seed_tail, ctx->seed_stage); } }
		else { if (ctx->seed_minb == 12) { KB_LAUNCH_SMEM((k_fm_seed_q<12, u64>), gq, KB_BLOCK, seed_smem, s, ixs, pm, bt, ctx->seed_qp, ctx->seed_qs, ctx->seed_trips > 0 ? ctx->seed_trips : 6, ctx->seed_tail, ctx->seed_stage); } else { KB_LAUNCH_SMEM((k_fm_seed_q<10, u64>), gq, KB_BLOCK, seed_smem, s, ixs, pm, bt, ctx->seed_qp, ctx->seed_qs, ctx->seed_trips > 0 ? ctx->seed_trips : 6, ctx->seed_tail, ctx->seed_stage); } }
	}
	else
	if (ctx->row32)
	{
		if (ctx->seed_minb == 8) { KB_LAUNCH((k_fm_seed<8, u32>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
		else if (ctx->seed_minb == 12) { KB_LAUNCH((k_fm_seed<12, u32>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
		else { KB_LAUNCH((k_fm_seed<10, u32>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
	}
	else
	{
		if (ctx->seed_minb == 8) { KB_LAUNCH((k_fm_seed<8, u64>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
		else if (ctx->seed_minb == 12) { KB_LAUNCH((k_fm_seed<12, u64>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
		else { KB_LAUNCH((k_fm_seed<10, u64>), g_reads, KB_BLOCK, s, ixs, pm, bt); }
	}
	sl.launches++;
	CK(cudaEventRecord(sl.ev[1], s));
	if (ix.sa_full != nullptr && !pm.pacbio) { KB_LAUNCH(k_sa_locate_reads, g_reads, KB_BLOCK, s, ix, bt); } else { KB_LAUNCH(k_sa_locate, g_hits, KB_BLOCK, s, ix, bt); }
	sl.launches++;
	CK(cudaEventRecord(sl.ev[2], s));
	KB_LAUNCH(k_cand_pair, g_items, KB_BLOCK, s, ix, pm, bt, ctx->cand_heavy); sl.launches++;
	if (ctx->cand_heavy && !pm.pacbio) { KB_LAUNCH(k_cand_heavy, 148 * 8, KB_BLOCK, s, ix, pm, bt); KB_LAUNCH(k_cand_heavy_finish, 148 * 8, KB_BLOCK, s, ix, pm, bt, ctx->cand_wide); sl.launches += 2; }
	if (pm.pacbio)
	{
		if (ctx->cand_heavy) { KB_LAUNCH(k_cand_pacbio_sort, 148 * 8, KB_BLOCK, s, bt); sl.launches++; }
		KB_LAUNCH(k_cand_pacbio, g_slow, KB_BLOCK, s, ix, pm, bt, ctx->cand_heavy); sl.launches++;
	}
	CK(cudaEventRecord(sl.ev[3], s));
	if (pm.paired)
	{
		unsigned rt = (unsigned)ctx->rescue_threads, g = (unsigned)bt.scratch_threads / rt; if (g > 148u * 32u) g = 148u * 32u;
		const unsigned g_jobs = g_items < 148u * 16u ? g_items : 148u * 16u;   // thread per rescue job, each a chain of dependent loads: as many in flight as there may be jobs
		KB_LAUNCH(k_rescue_plan, g_jobs, KB_BLOCK, s, ix, pm, bt); sl.launches++;
		if (ctx->rescue_fast)
		{
#ifndef KB_EMUL
			k_rescue_fast<<<148 * 8, KB_BLOCK, (KB_BLOCK / 32) * sizeof(KbRescueFast), s>>>(ix, pm, bt);
#else
			KB_LAUNCH(k_rescue_fast, 1, KB_BLOCK, s, ix, pm, bt);
#endif
			sl.launches++;
		}
		else { KB_LAUNCH(k_rescue_all_slow, 148 * 4, KB_BLOCK, s, bt); sl.launches++; }
		KB_LAUNCH(k_rescue_win, g, rt, s, ix, pm, bt); sl.launches++;
		KB_LAUNCH(k_rescue_commit, g_jobs, KB_BLOCK, s, ix, pm, bt); sl.launches++;
	}
	CK(cudaEventRecord(sl.ev[4], s));
	KB_LAUNCH(k_segments, g_reads, KB_BLOCK, s, ix, pm, bt, ctx->seg_slab); sl.launches++;
	KB_LAUNCH(k_segments_slow, g_slow, KB_BLOCK, s, ix, pm, bt); sl.launches++;
	CK(cudaEventRecord(sl.ev[5], s));
	{ int rc = launch_phase_b(ctx, sl); if (rc) return rc; }
	CK(cudaEventRecord(sl.ev[6], s));
	KB_LAUNCH(k_assemble, g_reads, KB_BLOCK, s, ix, pm, bt); sl.launches++;
	KB_LAUNCH(k_assemble_slow, g_slow, KB_BLOCK, s, ix, pm, bt); sl.launches++;
	CK(cudaEventRecord(sl.ev[7], s));
	KB_LAUNCH(k_finalize, g_items, KB_BLOCK, s, ix, pm, bt, sl.aln.p); sl.launches++;
	CK(cudaEventRecord(sl.ev[8], s));
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(sl.counters_host, sl.counters.p, KB_NCOUNTERS * sizeof(u32), cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(sl.work_dev_host, sl.work.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s));
	return KB_OK;
}

static void grow_factors(kb_ctx* ctx, u32 st)
{
	if (st & (KB_OVF_SEEDS | KB_OVF_CANDS | KB_OVF_HITS)) ctx->seg_factor *= 4;
	if (st & KB_OVF_CIGAR) ctx->cigar_factor *= 4;
	if (st & (KB_OVF_SCRATCH | KB_OVF_NW)) ctx->scratch_factor *= 2;
	if (st & KB_OVF_RESCUE) ctx->rtask_factor *= 8;
	if (st & KB_OVF_SEGX) ctx->segx_factor *= 4;
	if (st & KB_OVF_JOBS) ctx->job_factor *= 4;
	if (st & KB_OVF_RUNS) ctx->run_factor *= 4;
	if (st & KB_OVF_EXTRA) ctx->extra_factor *= 8;
}

// adds a finished slot's instrumentation to the context totals
static void account_slot(kb_ctx* ctx, kb_slot& sl)
{
	for (int i = 0; i < 8; i++) { float ms = 0; cudaEventElapsedTime(&ms, sl.ev[i], sl.ev[i + 1]); ctx->stage_ms[i] += ms; }
	{ float ms = 0; cudaEventElapsedTime(&ms, sl.ev[0], sl.ev[8]); ctx->stage_ms[8] += ms; }
	{ float ms = 0; cudaEventElapsedTime(&ms, sl.nw0, sl.nw1); ctx->stage_ms[9] += ms; }
	const unsigned long long* w = sl.work_dev_host; const u32* c = sl.counters_host;
	ctx->work_host[0] += w[0]; ctx->work_host[1] += w[1]; ctx->work_host[2] += w[2]; ctx->work_host[3] += w[3];
	ctx->work_host[4] += c[0]; ctx->work_host[5] += w[4]; ctx->work_host[6] += c[7]; ctx->work_host[7] += (uint64_t)sl.launches;
}

int kb_run(kb_ctx_t* ctx)
{
	if (!ctx) return KB_EINVAL;
	if (!ctx->staged) return fail(ctx, KB_ESTATE, "kb_run: no staged reads");
	CK(cudaSetDevice(ctx->device));
	kb_slot& sl = ctx->slot[0];
	memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms)); memset(ctx->work_host, 0, sizeof(ctx->work_host));
	ctx->ran_pipelined = false;
	if (sl.n_reads == 0) { ctx->ran = true; memset(ctx->counters_host, 0, sizeof(ctx->counters_host)); ctx->n_cigar_last = 0; return KB_OK; }
	for (int attempt = 0; attempt < 6; attempt++)
	{
		int rc = alloc_batch(ctx, sl, 0); if (rc) return rc;
		rc = launch_pipeline(ctx, sl); if (rc) return rc;
		CK(cudaStreamSynchronize(sl.stream));
		u32 st = sl.counters_host[3];
		if (st == 0)
		{
			account_slot(ctx, sl);
			memcpy(ctx->counters_host, sl.counters_host, sizeof(ctx->counters_host));
			ctx->n_cigar_last = sl.counters_host[2];
			ctx->ran = true;
			return KB_OK;
		}
		grow_factors(ctx, st);   // something overflowed: grow what was flagged and run the batch again
	}
	return fail(ctx, KB_EOVERFLOW, "kb_run: arenas still overflow after regrowth");
}

int kb_fetch_results(kb_ctx_t* ctx, kb_results_t* out)
{
	if (!ctx || !out) return KB_EINVAL;
	if (!ctx->ran) return fail(ctx, KB_ESTATE, "kb_fetch_results: nothing has run");
	CK(cudaSetDevice(ctx->device));
	kb_slot& sl = ctx->slot[0];
	if (ctx->ran_pipelined)   // kb_map_chunk delivered aln/pairs already; KB_ECAPACITY left the cigar elements in the chunk-wide arena
	{
		out->n_cigar = ctx->n_cigar_last;
		if (out->n_cigar > out->cap_cigar || (out->n_cigar > 0 && !out->cigar)) return fail(ctx, KB_ECAPACITY, "kb_fetch_results: cigar buffer too small");
		if (out->n_cigar) { CK(cudaMemcpyAsync(out->cigar, ctx->chunk_cigar.p, (size_t)out->n_cigar * 4, cudaMemcpyDeviceToHost, sl.stream)); CK(cudaStreamSynchronize(sl.stream)); }
		return KB_OK;
	}
	size_t n = (size_t)sl.n_reads;
	out->n_cigar = n ? ctx->n_cigar_last : 0;
	if (n == 0) return KB_OK;
	if (!out->aln || (out->n_cigar > 0 && !out->cigar)) return fail(ctx, KB_EINVAL, "kb_fetch_results: missing buffers");
	if (out->n_cigar > out->cap_cigar) return fail(ctx, KB_ECAPACITY, "kb_fetch_results: cigar buffer too small");
	CK(cudaMemcpyAsync(out->aln, sl.aln.p, n * sizeof(kb_aln_t), cudaMemcpyDeviceToHost, sl.stream));
	if (out->n_cigar) CK(cudaMemcpyAsync(out->cigar, sl.cigar.p, (size_t)out->n_cigar * 4, cudaMemcpyDeviceToHost, sl.stream));
	if (ctx->pm.paired && out->pairs) CK(cudaMemcpyAsync(out->pairs, sl.pstat.p, (n / 2) * sizeof(kb_pair_stat_t), cudaMemcpyDeviceToHost, sl.stream));
	CK(cudaStreamSynchronize(sl.stream));
	return KB_OK;
}

int kb_fetch_extra(kb_ctx_t* ctx, kb_extra_t* out, uint32_t cap, uint32_t* n)
{
	if (!ctx || !n) return KB_EINVAL;
	if (!ctx->ran || ctx->ran_pipelined) return fail(ctx, KB_ESTATE, "kb_fetch_extra: nothing has run");
	CK(cudaSetDevice(ctx->device));
	kb_slot& sl = ctx->slot[0];
	*n = (ctx->pm.multihit && sl.n_reads) ? ctx->counters_host[14] : 0;
	if (*n == 0) return KB_OK;
	if (!out || *n > cap) return fail(ctx, KB_ECAPACITY, "kb_fetch_extra: buffer too small");
	CK(cudaMemcpyAsync(out, sl.extra.p, (size_t)*n * sizeof(kb_extra_t), cudaMemcpyDeviceToHost, sl.stream));
	CK(cudaStreamSynchronize(sl.stream));
	std::sort(out, out + *n, [](const kb_extra_t& a, const kb_extra_t& b) { return a.read != b.read ? a.read < b.read : a.rank < b.rank; });
	return KB_OK;
}

// Large chunks: sub-batches rotate through the slots. Each slot's stream carries H2D -> kernels -> D2H of its sub-batch, so
// the copies of one sub-batch overlap the kernels of its neighbours. Cigar elements of all sub-batches go to one chunk-wide
// arena behind one cursor (offsets in kb_aln_t are chunk-wide). Only the assemble kernels hand out elements, so the cursor
// values on either side of them (counters[30], [31]) bracket a sub-batch's elements; when a sub-batch retires that range is
// copied back on copy_stream. Ranges of neighbouring sub-batches may overlap when their assemble phases ran side by side: the
// later copy (queued after its sub-batch finished, same stream) then rewrites those elements with their final values.
//
// The plan: the GPU idles while the first sub-batch is copied in and the copy engine alone is busy while the last one is
// copied out, so sub-batches start small (pipe_first), grow by pipe_grow percent up to the steady size, and the tail halves
// the remainder down to pipe_tail. Default (r14 sweep at C2, 2 M reads): steady = n/4, first = steady/2, no tail: 11.5 ms against
// 11.9 ms uniform n/4, 16.5 ms uniform n/10, 17.4 ms unpipelined. KB_PIPE_FIRST=-1 gives uniform sub-batches.
static void pipeline_plan(const kb_ctx* ctx, int n, std::vector<int>& first, std::vector<int>& count)
{
	int steady = ctx->pipe_sub_reads;
	if (steady <= 0) { steady = n / 4; if (steady < 65536) steady = 65536; if (steady > 1048576) steady = 1048576; }
	double s = ctx->pipe_first > 0 ? (double)ctx->pipe_first : (ctx->pipe_first == 0 ? (double)(steady / 2) : (double)steady);   // -1: uniform
	first.clear(); count.clear();
	for (int at = 0; at < n;)
	{
		int rem = n - at, c = s < (double)steady ? (int)s : steady;
		if (ctx->pipe_tail > 0) { int half = rem / 2 > ctx->pipe_tail ? rem / 2 : ctx->pipe_tail; if (c > half) c = half; }
		c &= ~1; if (c < 2) c = 2; if (c > rem) c = rem;
		if (rem - c < c / 4) c = rem;   // no crumbs
		first.push_back(at); count.push_back(c); at += c;
		s = s * (double)ctx->pipe_grow / 100.0;
	}
}

// an error in the middle of a pipelined chunk must not return while copies into the caller's buffers are still in flight
static int drain_pipeline(kb_ctx* ctx, int rc)
{
	for (int k = 0; k < KB_SLOTS; k++) if (ctx->slot[k].stream) cudaStreamSynchronize(ctx->slot[k].stream);
	if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
	cudaGetLastError();
	return rc;
}
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return drain_pipeline(ctx, fail(ctx, e_ == cudaErrorMemoryAllocation ? KB_ENOMEM : KB_ECUDA, #call, e_)); } while (0)

static int map_chunk_pipelined(kb_ctx* ctx, const KbReadSrc& in, const int32_t* est, kb_results_t* out)
{
	const int n = in.n;
	std::vector<int> p_first, p_count; pipeline_plan(ctx, n, p_first, p_count);
	const int nsub = (int)p_first.size();
	memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms)); memset(ctx->work_host, 0, sizeof(ctx->work_host));
	for (int attempt = 0; attempt < 6; attempt++)
	{
		size_t cap = (size_t)(ctx->cigar_factor * (double)n) + (ctx->pm.pacbio ? in.seq_off[n] / 2 : 0) + 65536;
		if (cap > 0xF0000000ull) cap = 0xF0000000ull;
		CK(cudaStreamSynchronize(ctx->copy_stream));   // copies of an abandoned attempt
		CK(ctx->chunk_cigar.ensure(cap)); CK(ctx->chunk_cursor.ensure(4));
		CK(cudaMemsetAsync(ctx->chunk_cursor.p, 0, 4 * sizeof(u32), ctx->slot[0].stream));
		CK(cudaEventRecord(ctx->chunk_start, ctx->slot[0].stream));
		CK(cudaStreamSynchronize(ctx->slot[0].stream));
		u32 status = 0, n_cigar = 0; bool fits = true;
		memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms)); memset(ctx->work_host, 0, sizeof(ctx->work_host));
		for (int k = 0; k < nsub + KB_SLOTS; k++)
		{
			kb_slot& sl = ctx->slot[k % KB_SLOTS];
			if (k >= KB_SLOTS)   // retire the sub-batch that used this slot
			{
				CKP(cudaStreamSynchronize(sl.stream));
				status |= sl.counters_host[3];
				account_slot(ctx, sl);
				const u32 lo = sl.counters_host[30], hi = sl.counters_host[31];
				if (hi > n_cigar) n_cigar = hi;
				if (hi > out->cap_cigar) fits = false;
				if (!status && fits && hi > lo) CKP(cudaMemcpyAsync(out->cigar + lo, ctx->chunk_cigar.p + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
				if (ctx->trace)
				{
					float a = 0, b = 0, c = 0, d = 0;
					cudaEventElapsedTime(&a, ctx->chunk_start, sl.ev[9]); cudaEventElapsedTime(&b, ctx->chunk_start, sl.ev[0]);
					cudaEventElapsedTime(&c, ctx->chunk_start, sl.ev[8]); cudaEventElapsedTime(&d, ctx->chunk_start, sl.done);
					fprintf(stderr, "[kb pipe] sub %d (%d reads): h2d %.2f..%.2f  kernels ..%.2f  d2h ..%.2f ms  cigar [%u, %u)\n", k - KB_SLOTS, sl.n_reads, a, b, c, d, lo, hi);
				}
			}
			if (status || k >= nsub) continue;
			const int first = p_first[k], count = p_count[k];
			CKP(cudaEventRecord(sl.ev[9], sl.stream));
			int rc = stage_slot(ctx, sl, in, first, count, est); if (rc) return drain_pipeline(ctx, rc);
			rc = alloc_batch(ctx, sl, 1); if (rc) return drain_pipeline(ctx, rc);
			rc = launch_pipeline(ctx, sl); if (rc) return drain_pipeline(ctx, rc);
			CKP(cudaMemcpyAsync(out->aln + first, sl.aln.p, (size_t)count * sizeof(kb_aln_t), cudaMemcpyDeviceToHost, sl.stream));
			if (ctx->pm.paired && out->pairs) CKP(cudaMemcpyAsync(out->pairs + first / 2, sl.pstat.p, (size_t)(count / 2) * sizeof(kb_pair_stat_t), cudaMemcpyDeviceToHost, sl.stream));
			CKP(cudaEventRecord(sl.done, sl.stream));
		}
		if (status == 0)
		{
			CK(cudaStreamSynchronize(ctx->copy_stream));
			out->n_cigar = n_cigar; ctx->n_cigar_last = n_cigar;
			ctx->ran = true; ctx->ran_pipelined = true;
			if (!fits) return fail(ctx, KB_ECAPACITY, "kb_map_chunk: cigar buffer too small");   // kb_fetch_results() with a larger buffer delivers them
			return KB_OK;
		}
		grow_factors(ctx, status);
	}
	return fail(ctx, KB_EOVERFLOW, "kb_map_chunk: arenas still overflow after regrowth");
}

static int map_chunk(kb_ctx* ctx, const KbReadSrc& in, const int32_t* est, kb_results_t* out, const char* who)
{
	for (int k = 0; k < KB_SLOTS; k++) if (ctx->slot[k].busy) return fail(ctx, KB_ESTATE, "a chunk is in flight (kb_map_chunk_end it first)");
	if (ctx->have_index && !ctx->pm.multihit && in.n >= ctx->pipe_min_reads && (in.seq || (in.pk && in.pk->code && (!in.pk->n_exc || in.pk->exc))) && in.seq_off && out->aln && out->cigar
	    && !(ctx->pm.paired && ((in.n & 1) || !est)))
	{
		CK(cudaSetDevice(ctx->device));
		ctx->staged = false; ctx->ran = false;
		return map_chunk_pipelined(ctx, in, est, out);
	}
	int rc = stage_reads(ctx, in, est, who); if (rc) return rc;
	rc = kb_run(ctx); if (rc) return rc;
	return kb_fetch_results(ctx, out);
}
int kb_map_chunk(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est, kb_results_t* out)
{
	if (!ctx || !in || !out) return fail(ctx, KB_EINVAL, "kb_map_chunk: bad arguments");
	return map_chunk(ctx, read_src(in), est, out, "kb_map_chunk: bad arguments");
}
int kb_map_chunk_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est, kb_results_t* out)
{
	if (!ctx || !in || !out) return fail(ctx, KB_EINVAL, "kb_map_chunk_packed: bad arguments");
	return map_chunk(ctx, read_src_packed(in), est, out, "kb_map_chunk_packed: bad arguments");
}

// ---- chunks in flight -------------------------------------------------------------------------------------------------
// kb_map_chunk overlaps copies and kernels INSIDE one chunk by cutting it into sub-batches; the price is kernel efficiency (a
// sub-batch's 22-kernel chain has a latency floor of ~1.5 ms, and small grids leave SMs idle: r19, C3, 2.5 M reads: 21.6 ms per chunk
// against 13.9 ms of kernels for the same reads as one batch). A caller that has the next chunk ready -- a mapper streaming a
// FASTQ file does -- gets the overlap ACROSS chunks instead: every chunk runs as one batch on its own slot and stream, the
// H2D copy of chunk k+1 and the D2H copy of chunk k-1 run under the kernels of chunk k. Up to KB_SLOTS chunks may be in flight.
// begin: returns at once with a ticket; end: waits for that chunk, delivers n_cigar and the cigar elements.
static int begin_slot(kb_ctx* ctx, kb_slot& sl, const KbReadSrc& in, const int32_t* est, kb_results_t* out)
{
	int rc = stage_slot(ctx, sl, in, 0, in.n, est); if (rc) return rc;
	if (in.n == 0) return KB_OK;
	rc = alloc_batch(ctx, sl, 0); if (rc) return rc;
	rc = launch_pipeline(ctx, sl); if (rc) return rc;
	CK(cudaMemcpyAsync(out->aln, sl.aln.p, (size_t)in.n * sizeof(kb_aln_t), cudaMemcpyDeviceToHost, sl.stream));
	if (ctx->pm.paired && out->pairs) CK(cudaMemcpyAsync(out->pairs, sl.pstat.p, (size_t)(in.n / 2) * sizeof(kb_pair_stat_t), cudaMemcpyDeviceToHost, sl.stream));
	return KB_OK;
}
static int map_chunk_begin(kb_ctx* ctx, const KbReadSrc& in, const int32_t* est, kb_results_t* out, int* ticket, const char* who)
{
	if (!ticket || in.n < 0 || (in.n > 0 && ((!in.seq && !in.pk) || !in.seq_off || !out->aln || !out->cigar))) return fail(ctx, KB_EINVAL, who);
	if (!ctx->have_index) return fail(ctx, KB_ENOINDEX, "kb_map_chunk_begin: no index");
	if (ctx->pm.multihit) return fail(ctx, KB_EINVAL, "kb_map_chunk_begin: not available with -m (use kb_map_chunk)");
	if (ctx->pm.paired && ((in.n & 1) || (in.n > 0 && !est))) return fail(ctx, KB_EINVAL, "kb_map_chunk_begin: paired chunks need an even read count and one EstDistance per pair");
	CK(cudaSetDevice(ctx->device));
	int k = -1;
	for (int i = 0; i < KB_SLOTS; i++) { const int c = (ctx->next_slot + i) % KB_SLOTS; if (!ctx->slot[c].busy) { k = c; break; } }
	if (k < 0) return fail(ctx, KB_ESTATE, "kb_map_chunk_begin: every slot holds a chunk in flight (call kb_map_chunk_end first)");
	kb_slot& sl = ctx->slot[k];
	ctx->staged = false; ctx->ran = false; ctx->ran_pipelined = false;
	int rc = begin_slot(ctx, sl, in, est, out);
	if (rc) { cudaStreamSynchronize(sl.stream); cudaGetLastError(); return rc; }
	sl.busy = true; sl.user_out = out; sl.user_est = est; sl.user_packed = in.pk != nullptr;
	if (in.pk) sl.user_pk = *in.pk; else { sl.user_text.n_reads = in.n; sl.user_text.seq = in.seq; sl.user_text.seq_off = in.seq_off; }
	ctx->next_slot = (k + 1) % KB_SLOTS;
	*ticket = k;
	return KB_OK;
}
int kb_map_chunk_begin(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est, kb_results_t* out, int* ticket)
{
	if (!ctx || !in || !out) return fail(ctx, KB_EINVAL, "kb_map_chunk_begin: bad arguments");
	return map_chunk_begin(ctx, read_src(in), est, out, ticket, "kb_map_chunk_begin: bad arguments");
}
int kb_map_chunk_begin_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est, kb_results_t* out, int* ticket)
{
	if (!ctx || !in || !out) return fail(ctx, KB_EINVAL, "kb_map_chunk_begin_packed: bad arguments");
	return map_chunk_begin(ctx, read_src_packed(in), est, out, ticket, "kb_map_chunk_begin_packed: bad arguments");
}
int kb_map_chunk_end(kb_ctx_t* ctx, int ticket)
{
	if (!ctx || ticket < 0 || ticket >= KB_SLOTS || !ctx->slot[ticket].busy) return fail(ctx, KB_ESTATE, "kb_map_chunk_end: no such chunk in flight");
	CK(cudaSetDevice(ctx->device));
	kb_slot& sl = ctx->slot[ticket]; kb_results_t* out = sl.user_out;
	sl.busy = false;
	memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms)); memset(ctx->work_host, 0, sizeof(ctx->work_host));
	if (sl.n_reads == 0) { out->n_cigar = 0; return KB_OK; }
	for (int attempt = 0; attempt < 6; attempt++)
	{
		CK(cudaStreamSynchronize(sl.stream));
		const u32 st = sl.counters_host[3];
		if (st == 0)
		{
			account_slot(ctx, sl);
			memcpy(ctx->counters_host, sl.counters_host, sizeof(ctx->counters_host));
			out->n_cigar = sl.counters_host[2]; ctx->n_cigar_last = out->n_cigar;
			if (out->n_cigar > out->cap_cigar) return fail(ctx, KB_ECAPACITY, "kb_map_chunk_end: cigar buffer too small");
			if (out->n_cigar) { CK(cudaMemcpyAsync(out->cigar, sl.cigar.p, (size_t)out->n_cigar * 4, cudaMemcpyDeviceToHost, sl.stream)); CK(cudaStreamSynchronize(sl.stream)); }
			return KB_OK;
		}
		grow_factors(ctx, st);   // something overflowed: grow what was flagged and run this chunk again on its slot
		const KbReadSrc in = sl.user_packed ? read_src_packed(&sl.user_pk) : read_src(&sl.user_text);
		int rc = begin_slot(ctx, sl, in, sl.user_est, out); if (rc) return rc;
	}
	return fail(ctx, KB_EOVERFLOW, "kb_map_chunk_end: arenas still overflow after regrowth");
}

uint64_t kb_packed_words(const kb_reads_t* in) { return (in && in->n_reads > 0 && in->seq_off) ? (in->seq_off[in->n_reads] >> 5) + (uint64_t)in->n_reads + 1 : 1; }
int kb_pack_reads(const kb_reads_t* in, uint64_t* code, uint64_t* exc, uint64_t cap_exc, int threads, kb_reads_packed_t* out)
{
	if (!in || !out || in->n_reads < 0 || (in->n_reads > 0 && (!in->seq || !in->seq_off || !code))) return KB_EINVAL;
	const int n = in->n_reads; const int T = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
	std::vector<std::vector<u64>> ex((size_t)T);
	u8 lut[256];   // nt4 code in the low two bits (0 for no base), bit 2 = "not one of the upper-case letters ACGT"
	for (int c = 0; c < 256; c++) { const int k = kb_nt4((u8)c); lut[c] = (u8)((k & 3) | ((k > 3 || (c & 0x20)) ? 4 : 0)); if (k > 3) lut[c] = 4; }
	auto work = [&](int t) {
		const int lo = (int)((long long)n * t / T), hi = (int)((long long)n * (t + 1) / T);
		std::vector<u64>& e = ex[(size_t)t];
		for (int r = lo; r < hi; r++)
		{
			const u64 off = in->seq_off[r]; const u64 len = in->seq_off[r + 1] - off; const u8* s = in->seq + off;
			u64* w = code + (off >> 5) + (u64)r;
			for (u64 b = 0; b < len; b += 32)
			{
				// table-driven and branch-free over the 32 characters; a word that holds anything but upper-case bases is walked again for the list
				const int m = len - b < 32 ? (int)(len - b) : 32; u64 v = 0; unsigned odd = 0;
				for (int i = 0; i < m; i++) { const unsigned q = lut[s[b + i]]; v |= (u64)(q & 3u) << (62 - 2 * i); odd |= q; }
				w[b >> 5] = v;
				if (odd & 4u) for (int i = 0; i < m; i++) { const u8 c = s[b + i]; if (lut[c] & 4u) e.push_back(((u64)(u32)r << 32) | ((b + (u64)i) << 8) | c); }
			}
		}
	};
	if (T == 1) work(0);
	else { std::vector<std::thread> th; for (int t = 1; t < T; t++) th.emplace_back(work, t); work(0); for (auto& x : th) x.join(); }
	u64 total = 0; for (auto& e : ex) total += e.size();
	out->n_reads = n; out->code = code; out->n_words = kb_packed_words(in); out->seq_off = in->seq_off; out->exc = exc; out->n_exc = total;
	if (total > cap_exc || (total && !exc)) return KB_ECAPACITY;
	u64 at = 0; for (auto& e : ex) { if (!e.empty()) memcpy(exc + at, e.data(), e.size() * 8); at += e.size(); }
	return KB_OK;
}

int kb_stage_ms(kb_ctx_t* ctx, float* ms, int n) { if (!ctx || !ms) return KB_EINVAL; int k = n < 10 ? n : 10; for (int i = 0; i < k; i++) ms[i] = ctx->stage_ms[i]; return k; }
int kb_work(kb_ctx_t* ctx, uint64_t* w, int n) { if (!ctx || !w) return KB_EINVAL; int k = n < 8 ? n : 8; for (int i = 0; i < k; i++) w[i] = ctx->work_host[i]; return k; }

int64_t kb_debug_fetch(kb_ctx_t* ctx, int what, void* dst, uint64_t bytes)
{
	if (!ctx || !dst) return KB_EINVAL;
	if (!ctx->ran || ctx->ran_pipelined) return KB_ESTATE;
	cudaSetDevice(ctx->device);
	kb_slot& sl = ctx->slot[0];
	size_t n = (size_t)sl.n_reads; const void* src = nullptr; size_t have = 0;
	switch (what)
	{
	case 0: src = sl.n_seeds.p; have = n * 4; break;
	case 1: src = sl.seed_off.p; have = n * 4; break;
	case 2: src = sl.segs.p; have = (size_t)ctx->counters_host[0] * sizeof(KbSeg); break;
	case 3: src = sl.n_cands.p; have = n * 4; break;
	case 4: src = sl.cand_off.p; have = n * 4; break;
	case 5: src = sl.cands.p; have = (size_t)ctx->counters_host[1] * sizeof(KbCand); break;
	case 6: src = sl.reports.p; have = (size_t)ctx->counters_host[1] * sizeof(KbReport); break;
	case 7: src = sl.res.p; have = n * sizeof(KbReadRes); break;
	case 8: src = sl.cigar.p; have = (size_t)ctx->counters_host[2] * 4; break;
	case 9: src = sl.counters.p; have = KB_NCOUNTERS * 4; break;
	case 10: src = sl.jobs.p; have = (size_t)ctx->counters_host[11] * sizeof(KbJob); break;
	case 11: src = sl.hits.p; have = n * (size_t)sl.bt.max_hits * sizeof(KbHit); break;
	case 12: src = sl.n_hits.p; have = n * 4; break;
	case 13: { if (bytes < 4) return KB_EINVAL; const int32_t mh = sl.bt.max_hits; memcpy(dst, &mh, 4); return 4; }
	default: return KB_EINVAL;
	}
	if (have > bytes) have = bytes;
	if (have && cudaMemcpy(dst, src, have, cudaMemcpyDeviceToHost) != cudaSuccess) return KB_ECUDA;
	return (int64_t)have;
}

int kb_debug_align(kb_ctx_t* ctx, const kb_dbg_frag_t* specs, int n, kb_dbg_frag_out_t* out, uint32_t* ops, uint32_t cap_ops)
{
	if (!ctx || !specs || !out || !ops || n <= 0) return fail(ctx, KB_EINVAL, "kb_debug_align: bad arguments");
	if (!ctx->staged) return fail(ctx, KB_ESTATE, "kb_debug_align: no staged reads");
	CK(cudaSetDevice(ctx->device));
	kb_slot& sl = ctx->slot[0]; cudaStream_t s = sl.stream;
	int rc = alloc_batch(ctx, sl, 0); if (rc) return rc;
	KbBatchDev& bt = sl.bt;
	std::vector<kb_dbg_frag_out_t> res((size_t)n); u64 need_ops = 0, need_runs = 0;
	for (int i = 0; i < n; i++)
	{
		const kb_dbg_frag_t& f = specs[i];
		if (f.read >= (uint32_t)sl.n_reads || f.rpos < 0 || f.rlen < 0 || f.glen < 0 || f.mode < 0 || f.mode > 4) return fail(ctx, KB_EINVAL, "kb_debug_align: bad fragment");
		memset(&res[i], 0, sizeof(res[i])); res[i].ops_off = (uint32_t)need_ops;
		need_ops += (u64)(f.rlen + f.glen + 4); need_runs += (u64)(f.rlen + f.glen + 2);
	}
	if (need_ops > cap_ops || need_runs > sl.cap_runs || (size_t)n > sl.cap_jobs || (size_t)n > sl.cap_segx) return fail(ctx, KB_ECAPACITY, "kb_debug_align: too many / too large fragments for this batch's arenas");
	DevBuf<kb_dbg_frag_t> d_specs; DevBuf<kb_dbg_frag_out_t> d_res; DevBuf<u32> d_ops;
	CK(d_specs.ensure((size_t)n)); CK(d_res.ensure((size_t)n)); CK(d_ops.ensure((size_t)need_ops));
	CK(cudaMemcpyAsync(d_specs.p, specs, (size_t)n * sizeof(kb_dbg_frag_t), cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(d_res.p, res.data(), (size_t)n * sizeof(kb_dbg_frag_out_t), cudaMemcpyHostToDevice, s));
	CK(cudaMemsetAsync(sl.counters.p, 0, KB_NCOUNTERS * sizeof(u32), s)); CK(cudaMemsetAsync(sl.work.p, 0, 8 * sizeof(u64), s));
	sl.launches = 0;
	launch_pack(sl);
	KB_LAUNCH(k_debug_classify, (unsigned)((n + KB_BLOCK - 1) / KB_BLOCK), KB_BLOCK, s, ctx->ix, ctx->pm, bt, d_specs.p, n);
	rc = launch_phase_b(ctx, sl);
	if (rc == KB_OK)
	{
		KB_LAUNCH(k_debug_assemble, (unsigned)((n + KB_BLOCK - 1) / KB_BLOCK), KB_BLOCK, s, ctx->ix, ctx->pm, bt, d_specs.p, n, d_res.p, d_ops.p);
		cudaError_t e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_res.p, (size_t)n * sizeof(kb_dbg_frag_out_t), cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(ops, d_ops.p, (size_t)need_ops * 4, cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(sl.counters_host, sl.counters.p, KB_NCOUNTERS * sizeof(u32), cudaMemcpyDeviceToHost, s);
		if (e != cudaSuccess) rc = fail(ctx, KB_ECUDA, "kb_debug_align", e);
	}
	cudaError_t e2 = cudaStreamSynchronize(s);
	d_specs.release(); d_res.release(); d_ops.release();
	if (rc) return rc;
	if (e2 != cudaSuccess) return fail(ctx, KB_ECUDA, "kb_debug_align: kernels", e2);
	if (sl.counters_host[3]) return fail(ctx, KB_EOVERFLOW, "kb_debug_align: a device arena overflowed");
	memcpy(ctx->counters_host, sl.counters_host, sizeof(ctx->counters_host));
	ctx->ran = true; ctx->ran_pipelined = false;   // kb_debug_fetch may read the arenas (counters, jobs, ...) of this run
	return KB_OK;
}

#ifdef KB_EMUL   // the device index builder (kb_index_build.cu) has no host emulation: the emulated library reports that there is no device
int kb_index_build(int, const uint8_t*, int64_t, kb_built_index_t* out) { if (out) memset(out, 0, sizeof(*out)); return KB_ENODEV; }
void kb_index_free(kb_built_index_t*) {}
const char* kb_index_build_error(void) { return "no CUDA device (emulation build)"; }
#endif

} // extern "C"
