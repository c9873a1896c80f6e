// FM-index primitives on the re-blocked 32-byte Occ layout, MEM seeding and SA locate.
// Replaces (reference src/bwt_search.cpp): bwt_occ4 :68, bwt_2occ4 :87, bwt_occ :44, bwt_invPsi :120,
// bwt_sa :128, BWT_Search :140; and the seeding drivers IdentifySeedPairs_FastMode / _SensitiveMode
// (src/AlignmentCandidates.cpp:49,132).
//
// The per-item logic is written as KB_HD functions so that tests/emul can compile the very same source for the
// host and step through it next to the oracle. The product only ever runs the __global__ kernels (kb_kernels.cu).
#ifndef KB_FM_CUH
#define KB_FM_CUH
#include "kb_types.h"

#if defined(__CUDACC__) && !defined(KB_EMUL)
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cooperative_groups/scan.h>
#endif

#if defined(__CUDA_ARCH__)
#define KB_CLZLL(x) __clzll((long long)(x))
#define KB_CLZ(x) __clz((int)(x))
#define KB_POPCLL(x) __popcll(x)
#define KB_LDG4(p) __ldg(p)
#define KB_LDG(p) __ldg(p)
#define KB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define KB_ATOMIC_OR(p, v) atomicOr((p), (v))
#define KB_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define KB_ATOMIC_CAS(p, c, v) atomicCAS((p), (c), (v))
#define KB_ATOMIC_EXCH(p, v) atomicExch((p), (v))
#else
#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; };
#endif
#define KB_POPCLL(x) __builtin_popcountll(x)
#define KB_CLZLL(x) ((x) ? __builtin_clzll((unsigned long long)(x)) : 64)
#define KB_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#define KB_LDG4(p) (*(p))
#define KB_LDG(p) (*(p))
template <class T, class V> static inline T kb_host_add(T* p, V v) { T o = *p; *p = (T)(o + v); return o; }
template <class T, class V> static inline T kb_host_or(T* p, V v) { T o = *p; *p = (T)(o | v); return o; }
template <class T, class V> static inline T kb_host_max(T* p, V v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
#define KB_ATOMIC_ADD(p, v) kb_host_add((p), (v))
#define KB_ATOMIC_OR(p, v) kb_host_or((p), (v))
template <class T> static inline T kb_host_cas(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> static inline T kb_host_exch(T* p, T v) { T o = *p; *p = v; return o; }
#define KB_ATOMIC_MAX(p, v) kb_host_max((p), (v))
#define KB_ATOMIC_CAS(p, c, v) kb_host_cas((p), (c), (v))
#define KB_ATOMIC_EXCH(p, v) kb_host_exch((p), (v))
#endif

// ---- bump allocation from a grid-wide cursor ----------------------------------------------------------------------
// Every per-batch arena (seeds, candidates, segments, jobs, pieces, cigar elements) is handed out by one 32-bit cursor in
// HBM. One atomicAdd per item makes that cursor the bottleneck of whole kernels: same-address atomics retire at roughly one
// per nanosecond at the L2, which for 2 M reads is milliseconds (ncu, profiles/r11: half of k_segments' stall samples sat
// on three such atomics). So the lanes of a warp that arrive at an allocation together (whatever subset that is) combine
// their requests: exclusive scan inside the coalesced group, ONE atomicAdd by its first lane, base broadcast back.
// Which slots a lane gets depends on scheduling, exactly as with plain atomics; nothing downstream depends on slot order.
#if defined(__CUDA_ARCH__)
template <class T> __device__ __forceinline__ T kb_alloc_slots(T* cursor, T n)
{
	namespace cg = cooperative_groups;
	cg::coalesced_group g = cg::coalesced_threads();
	if (g.size() == 1) return atomicAdd(cursor, n);
	const T pre = cg::exclusive_scan(g, n);
	T base = 0;
	if (g.thread_rank() == g.size() - 1) base = atomicAdd(cursor, (T)(pre + n));
	return g.shfl(base, g.size() - 1) + pre;
}
// the same when lanes may name different cursors (one group per distinct cursor)
__device__ __forceinline__ u32 kb_alloc_slots_keyed(u32* cursor, u32 n)
{
	namespace cg = cooperative_groups;
	cg::coalesced_group a = cg::coalesced_threads();
	if (a.size() == 1) return atomicAdd(cursor, n);
	cg::coalesced_group g = cg::labeled_partition(a, (unsigned long long)cursor);
	const u32 pre = cg::exclusive_scan(g, n);
	u32 base = 0;
	if (g.thread_rank() == g.size() - 1) base = atomicAdd(cursor, pre + n);
	return g.shfl(base, g.size() - 1) + pre;
}
__device__ __forceinline__ void kb_max_u32(u32* dst, u32 v)
{
	namespace cg = cooperative_groups;
	cg::coalesced_group g = cg::coalesced_threads();
	const u32 m = cg::reduce(g, v, cg::greater<u32>());
	if (g.thread_rank() == 0 && m) atomicMax(dst, m);
}
#define KB_ALLOC(p, n) kb_alloc_slots((p), (n))
#define KB_ALLOC_KEYED(p, n) kb_alloc_slots_keyed((p), (n))
#define KB_MAX_U32(p, v) kb_max_u32((p), (v))
#else
#define KB_ALLOC(p, n) KB_ATOMIC_ADD((p), (n))
#define KB_ALLOC_KEYED(p, n) KB_ATOMIC_ADD((p), (n))
#define KB_MAX_U32(p, v) KB_ATOMIC_MAX((p), (v))
#endif

// nst_nt4_table (src/BWT_Index/bntseq.c:40): A/a 0, C/c 1, G/g 2, T/t 3, everything else 4
KB_HD int kb_nt4(u8 c)
{
	// Branch-free on purpose: a ternary chain here makes the compiler clone everything downstream per base value
	// (jump threading), which turns the data-dependent base into 4-way warp divergence (measured: 8.6 of 32 lanes active).
	// ASCII: A 0x41, C 0x43, G 0x47, T 0x54 (+0x20 for lower case): code = ((c>>1) ^ (c>>2)) & 3 ; letters sit at 1,3,7,20.
	u32 x = c;
	u32 code = ((x >> 1) ^ (x >> 2)) & 3u;
	u32 valid = ((x & 0xC0u) == 0x40u ? 1u : 0u) & (0x0010008Au >> (x & 31u));
	return (int)((code & (0u - valid)) | (4u & (valid - 1u)));
}

// ---- packed reads (k_pack) -----------------------------------------------------------------------
KB_HD const KbPk* kb_pk_read(const KbBatchDev& bt, int r) { return bt.pk + ((bt.seq_off[r] >> 5) + (u64)r); }
KB_HD KbPk kb_load_pk(const KbPk* p)
{
#if defined(__CUDA_ARCH__)
	uint4 v = *reinterpret_cast<const uint4*>(p);
	KbPk k; k.code = ((u64)v.y << 32) | v.x; k.n4 = v.z; k.bad = v.w; return k;
#else
	return *p;
#endif
}
// the same word, streamed: not kept in L1 (the seeding kernel copies a read's words to shared memory once)
KB_HD KbPk kb_load_pk_once(const KbPk* p)
{
#if defined(__CUDA_ARCH__)
	u32 a, b, c, d;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
	KbPk k; k.code = ((u64)b << 32) | a; k.n4 = c; k.bad = d; return k;
#else
	return *p;
#endif
}
// one thread per (read, word): 32 characters -> KbPk
KB_HD void kb_pack_word(const KbBatchDev& bt, int r, int w)
{
	const u64 off = bt.seq_off[r]; const int len = (int)(bt.seq_off[r + 1] - off);
	if (32 * w >= len) return;
	const u8* s = bt.seq + off + 32 * w; const int n = len - 32 * w < 32 ? len - 32 * w : 32;
	u64 code = 0; u32 n4 = n == 32 ? 0u : (~0u >> n), bad = n4;
	for (int i = 0; i < n; i++)
	{
		u32 c = s[i]; u32 v = (u32)kb_nt4((u8)c);
		code |= (u64)(v & 3u) << (62 - 2 * i);
		u32 inval = v >> 2;
		n4 |= inval << (31 - i); bad |= (inval | ((c >> 5) & 1u)) << (31 - i);
	}
	KbPk k; k.code = code; k.n4 = n4; k.bad = bad;
	bt.pk[(off >> 5) + (u64)r + (u64)w] = k;
}
// 32 characters of a read starting at pos (bits beyond the end of the read are flagged in n4/bad, or garbage beyond the last word)
KB_HD KbPk kb_read_win(const KbPk* rd, int pos)
{
	KbPk a = kb_load_pk(rd + (pos >> 5)); const int s = pos & 31;
	if (s == 0) return a;
	KbPk b = kb_load_pk(rd + (pos >> 5) + 1);
	a.code = (a.code << (2 * s)) | (b.code >> (64 - 2 * s));
	a.n4 = (a.n4 << s) | (b.n4 >> (32 - s)); a.bad = (a.bad << s) | (b.bad >> (32 - s));
	return a;
}

// number of A,C,G,T among the first n (1..32) symbols of a 64-bit word (symbol i at bits 62-2i)
KB_HD void kb_count32(u64 W, int n, u32 cnt[4])
{
	const u64 M5 = 0x5555555555555555ull;
	u64 mask = (n >= 32) ? M5 : (M5 & ~(~0ull >> (2 * n)));
	u64 lo = W & mask, hi = (W >> 1) & mask;
	u32 t = (u32)KB_POPCLL(hi & lo), g = (u32)KB_POPCLL(hi & ~lo), c = (u32)KB_POPCLL(lo & ~hi);
	cnt[3] += t; cnt[2] += g; cnt[1] += c; cnt[0] += (u32)n - t - g - c;
}

// ---- one 32-byte Occ block = one DRAM sector, fetched with a single 256-bit load (LDG.E.256 on sm_100a) ----
// [u32 cntA,cntC,cntG,cntT : occurrences before the block][u64 lo][u64 hi]: the 64 BWT symbols of the block as two bit
// planes (low and high bit of the 2-bit code), row i of the block at bit 63-i. Rank queries are then one mask per plane.
struct KbBlk { u32 c0, c1, c2, c3; u64 lo, hi; };
// hint (KbIndexDev::ld_hint, seeding only): 1 = the sector is not kept in L1 (ld.global.nc.L1::no_allocate). Occ blocks, seeding-table
// entries and SA entries are visited once, at random; allocating them evicts the packed read words the same lanes come back to.
KB_HD KbBlk kb_load_blk(const uint32_t* occ, u64 blk, int hint = 0)
{
	KbBlk b; const uint32_t* p = occ + (blk << 3);
#if defined(__CUDA_ARCH__)
	u32 l0, l1, h0, h1;
	if (hint & 1)
		asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(b.c0), "=r"(b.c1), "=r"(b.c2), "=r"(b.c3), "=r"(l0), "=r"(l1), "=r"(h0), "=r"(h1) : "l"(p));
	else
		asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(b.c0), "=r"(b.c1), "=r"(b.c2), "=r"(b.c3), "=r"(l0), "=r"(l1), "=r"(h0), "=r"(h1) : "l"(p));
	b.lo = ((u64)l1 << 32) | l0; b.hi = ((u64)h1 << 32) | h0;
#else
	(void)hint;
	b.c0 = p[0]; b.c1 = p[1]; b.c2 = p[2]; b.c3 = p[3]; b.lo = ((u64)p[5] << 32) | p[4]; b.hi = ((u64)p[7] << 32) | p[6];
#endif
	return b;
}
KB_HD u64 kb_load_u64_once(const u64* p, int hint)
{
#if defined(__CUDA_ARCH__)
	if (hint & 1) { u64 v; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }
#else
	(void)hint;
#endif
	return KB_LDG(p);
}
KB_HD u64 kb_top_bits(u32 n) { return n == 0 ? 0ull : (~0ull << (64u - n)); }   // the first n (0..64) rows of a block
KB_HD u32 kb_sel4(int b, u32 v0, u32 v1, u32 v2, u32 v3) { u32 lo = (b & 1) ? v1 : v0, hi = (b & 1) ? v3 : v2; return (b & 2) ? hi : lo; }

// Occ(row, b) and the sum over bases j > b of Occ(row, j), `off` = the row's offset in its block
KB_HD void kb_rank_eq_gt(const KbBlk& k, int off, int b, u32* eq, u32* gt)
{
	const u64 nBL = (b & 1) ? 0ull : ~0ull, nBH = (b & 2) ? 0ull : ~0ull;
	const u64 pm = kb_top_bits((u32)off + 1u);
	const u64 e = (k.lo ^ nBL) & (k.hi ^ nBH);                 // symbol == b
	const u64 t = k.lo & nBL;
	const u64 g = (k.hi & t) | (nBH & (k.hi | t));             // symbol > b
	*eq = kb_sel4(b, k.c0, k.c1, k.c2, k.c3) + (u32)KB_POPCLL(e & pm);
	*gt = kb_sel4(b, k.c1 + k.c2 + k.c3, k.c2 + k.c3, k.c3, 0u) + (u32)KB_POPCLL(g & pm);
}

// ---- one forward extension of a bi-interval by base c (BWT_Search :151-166 with bwt_2occ4 :87) -------------------------
// b = 3 - c is the base looked up in the BWT. Returns false (state untouched) when the extended interval is empty.
// Rows k'+1 .. l' usually lie in one 32-byte block (always for a one-row interval): then that block alone gives
// Occ(k',b) = count before the block + matches among the rows in front of k'+1, and Occ(l',.)-Occ(k',.) from a range mask.
// ROW is the integer type of BWT row numbers: u32 when the text (2G + 1 rows) fits 32 bits, which halves the integer work of a
// step on small genomes, where the kernel is ALU-bound (the index sits in L2); u64 otherwise (HBM-bound there anyway).
template <class ROW>
KB_HD bool kb_extend(const KbIndexDev& ix, ROW& x0, ROW& x1, ROW& x2, int c, u32* blocks)
{
	const ROW primary = (ROW)ix.primary;
	const int b = 3 - c;
	const ROW k = x1 - 1, l = k + x2;
	const ROW rk = k - (ROW)(k >= primary), rl = l - (ROW)(l >= primary);
	const u64 nBL = (b & 1) ? 0ull : ~0ull, nBH = (b & 2) ? 0ull : ~0ull;
	u32 ek, n2, gt;
	if (((rk + 1) >> 6) == (rl >> 6))
	{
		const KbBlk B = kb_load_blk(ix.occ, (u64)(rl >> 6), ix.ld_hint); *blocks += 1;
		const u64 pk = kb_top_bits((u32)((rk + 1) & 63)), rg = kb_top_bits((u32)(rl & 63) + 1u) & ~pk;   // rows (k', l']
		const u64 e = (B.lo ^ nBL) & (B.hi ^ nBH);
		n2 = (u32)KB_POPCLL(e & rg);
		if (n2 == 0) return false;
		ek = kb_sel4(b, B.c0, B.c1, B.c2, B.c3) + (u32)KB_POPCLL(e & pk);
		gt = 0;
		if (x2 > 1)   // a one-row interval that survives holds b itself: nothing greater
		{
			const u64 t = B.lo & nBL;
			gt = (u32)KB_POPCLL(((B.hi & t) | (nBH & (B.hi | t))) & rg);
		}
	}
	else
	{
		const KbBlk bk = kb_load_blk(ix.occ, (u64)(rk >> 6), ix.ld_hint), bl = kb_load_blk(ix.occ, (u64)(rl >> 6), ix.ld_hint); *blocks += 2;
		u32 gk, el, gl;
		kb_rank_eq_gt(bk, (int)(rk & 63), b, &ek, &gk);
		kb_rank_eq_gt(bl, (int)(rl & 63), b, &el, &gl);
		n2 = el - ek; gt = gl - gk;
		if (n2 == 0) return false;
	}
	x0 = x0 + (ROW)((x1 <= primary && x1 + x2 - 1 >= primary) ? 1 : 0) + (ROW)gt;
	x1 = (ROW)ix.L2[b] + 1 + (ROW)ek; x2 = (ROW)n2;
	return true;
}

// one entry of the seeding table: BWT_Search over the K bases spelled by `kmer` (first base most significant)
KB_HD KbKtab kb_ktab_pack(u64 x0, u64 x1, u64 x2, u32 flen)
{
	KbKtab e;
	if (x2 == 0) { e.a = 0; e.b = flen; return e; }
	e.a = x0 | ((x2 & 0xFFFFFFull) << 40); e.b = x1 | ((x2 >> 24) << 40);
	return e;
}
KB_HD KbKtabE kb_ktab_unpack(u64 a, u64 b)
{
	KbKtabE e; const u64 M40 = (1ull << 40) - 1;
	e.x2 = (a >> 40) | ((b >> 40) << 24);
	if (e.x2 == 0) { e.x0 = 0; e.x1 = 0; e.flen = (u32)b; return e; }
	e.x0 = a & M40; e.x1 = b & M40; e.flen = 0;
	return e;
}
KB_HD KbKtab kb_ktab_entry(const KbIndexDev& ix, u64 kmer, int K)
{
	int p = (int)((kmer >> (2 * (K - 1))) & 3u);
	u64 x0 = ix.L2[p] + 1, x1 = ix.L2[3 - p] + 1, x2 = ix.L2[p + 1] - ix.L2[p];
	u32 blocks = 0;
	for (int i = 1; i < K; i++)
	{
		int c = (int)((kmer >> (2 * (K - 1 - i))) & 3u);
		if (!kb_extend(ix, x0, x1, x2, c, &blocks)) return kb_ktab_pack(0, 0, 0, (u32)i);
	}
	if (x2 == 0) return kb_ktab_pack(0, 0, 0, (u32)K);   // a base that does not occur (K == 1 only)
	return kb_ktab_pack(x0, x1, x2, 0);
}
KB_HD KbKtabE kb_load_ktab(const KbKtab* p, int hint = 0)
{
#if defined(__CUDA_ARCH__)
	u32 r[4];
	if (hint & 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
	else asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
	return kb_ktab_unpack(((u64)r[1] << 32) | r[0], ((u64)r[3] << 32) | r[2]);
#else
	(void)hint;
	return kb_ktab_unpack(p->a, p->b);
#endif
}

// LF step (bwt_invPsi :120 with bwt_occ :44 folded in; one 32-byte block per step)
KB_HD u64 kb_lf(const KbIndexDev& ix, u64 k)
{
	if (k == ix.primary) return 0;
	u64 r = k - (k > ix.primary);
	KbBlk bk = kb_load_blk(ix.occ, r >> 6);
	int off = (int)(r & 63);
	int sym = (int)((bk.lo >> (63 - off)) & 1ull) | (int)(((bk.hi >> (63 - off)) & 1ull) << 1);
	u32 eq, gt; kb_rank_eq_gt(bk, off, sym, &eq, &gt);
	return ix.L2[sym] + eq;
}

// SA locate (bwt_sa :128). Returns the text position; *steps receives the number of LF steps walked.
KB_HD u64 kb_sa(const KbIndexDev& ix, u64 k, u32* steps)
{
	if (ix.sa_full != nullptr) { *steps = 0; return KB_LDG(ix.sa_full + k); }
	u64 s = 0, mask = (u64)ix.sa_intv - 1;
	while (k & mask) { s++; k = kb_lf(ix, k); }
	*steps = (u32)s;
	return s + KB_LDG(ix.sa + k / (u64)ix.sa_intv);
}

KB_HD u64 kb_ref_win(const KbIndexDev& ix, i64 p, u32* inval);   // kb_align.cuh: 32 characters of the 2G text as 2-bit codes

// A search whose interval has shrunk to ONE row has a single occurrence in the text, at SA[x0]: every further extension step
// only asks whether the next read base equals the next text base (the text is closed under reverse complement, so the forward
// extension of P by c succeeds iff P's occurrence is followed by c; it fails at the end of the text, where kb_extend finds the
// row next to `primary` empty; x0 of a surviving one-row interval never moves). With the full SA in HBM the rest of the search
// is therefore one 8-byte load and a 32-bases-per-word comparison against the packed reference instead of one Occ block per
// base: same length, same x0, same step count. Returns the number of bases matched; *fail = stopped by a base that does not
// extend (what the step loop counts as one more, failed, step), as opposed to the search limit or a non-ACGT read character.
KB_HD int kb_unique_tail(const KbIndexDev& ix, const KbPk* rd, u64 row, int done, int cur, int lim, bool* fail, u32* blocks)
{
	const u64 M5 = 0x5555555555555555ull;
	const i64 tp = (i64)kb_load_u64_once(ix.sa_full + row, ix.ld_hint) + (i64)done;   // text position facing read position cur
	*blocks += 1; *fail = false;
	const int total = lim - cur;
	if (total <= 0) return 0;
	// Nearly every comparison lies inside one strand of the 2G text. Then the windows are cut at the READ's word boundaries (no funnel
	// shift of code / n4 on the read side, one 16-byte load per window), the strand is decided once, and the reference side is the
	// two-word funnel shift alone (reverse strand: of the mirrored forward window, bit-reversed and complemented). ~25 instructions per
	// 32 bases instead of ~70 (r33: the pass around this function was the bulk of the kernel's 4750 instructions per read at 8 lanes).
	// The result does not depend on how the stretch is cut into windows; `blocks` counts 32-base windows from `cur` as before.
	const bool fwd = tp >= 0 && tp + (i64)total + 32 <= ix.G, rev = tp >= ix.G && tp + (i64)total + 32 <= ix.G2;
	if ((fwd || rev) && !(ix.ld_hint & 2))   // bit 1 of ld_hint: KB_SEED_TAIL_FAST=0, the general loop below for everything (A/B)
	{
		int pos = cur; bool stopped = false;
		while (pos < lim)
		{
			const int o = pos & 31;
			const KbPk rw = kb_load_pk(rd + (pos >> 5));
			const int room = 32 - o, left = lim - pos, want = left < room ? left : room;
			const u64 rc = rw.code << (2 * o); const u32 rn = rw.n4 << o;   // read bases pos.. at the top; what is shifted in lies beyond `want`
			const i64 q = tp + (i64)(pos - cur);
			u64 gw;
			if (fwd)
			{
				const u64* w = ix.ref64 + (q >> 5); const int sh = (int)(q & 31) * 2;
				gw = KB_LDG(w); if (sh) gw = (gw << sh) | (KB_LDG(w + 1) >> (64 - sh));
			}
			else
			{
				const i64 f = ix.G2 - 32 - q;
				const u64* w = ix.ref64 + (f >> 5); const int sh = (int)(f & 31) * 2;
				u64 v = KB_LDG(w); if (sh) v = (v << sh) | (KB_LDG(w + 1) >> (64 - sh));
#if defined(__CUDA_ARCH__)
				v = __brevll(v);
#else
				{ u64 x = v; x = ((x >> 1) & M5) | ((x & M5) << 1); x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
				  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4); v = __builtin_bswap64(x); }
#endif
				gw = ~(((v >> 1) & M5) | ((v & M5) << 1));
			}
			u64 x = rc ^ gw; x = (x | (x >> 1)) & M5;
			const int i_mis = x ? (int)KB_CLZLL(x) >> 1 : 32, i_n = rn ? (int)KB_CLZ(rn) : 32;
			const int st = i_mis < i_n ? i_mis : i_n;
			if (st >= want) { pos += want; continue; }
			pos += st; *fail = i_n != st; stopped = true;   // a non-ACGT read character ends the search without a step (it is looked at first)
			break;
		}
		const int m = pos - cur;
		*blocks += stopped ? (u32)(m >> 5) + 1u : (u32)((total + 31) >> 5);
		return m;
	}
	int m = 0;
	while (cur + m < lim)
	{
		const KbPk rw = kb_read_win(rd, cur + m); u32 ginv; const u64 gw = kb_ref_win(ix, tp + m, &ginv); *blocks += 1;
		const int want = lim - (cur + m) < 32 ? lim - (cur + m) : 32;
		u64 x = rw.code ^ gw; x = (x | (x >> 1)) & M5;
		const int i_mis = x ? (int)KB_CLZLL(x) >> 1 : 32, i_n = rw.n4 ? (int)KB_CLZ(rw.n4) : 32, i_t = ginv ? (int)KB_CLZ(ginv) : 32;
		int s = i_mis < i_t ? i_mis : i_t; if (i_n < s) s = i_n;
		if (s >= want) { m += want; continue; }
		m += s; *fail = i_n != s;   // a non-ACGT read character ends the search without a step (it is looked at first)
		break;
	}
	return m;
}
// The same for a FEW rows (2 .. KB_SEED_TAIL_MAX occurrences: multi-copy genes, insertion sequences, diverged repeat copies):
// the walk would go on until no occurrence extends, i.e. for max_i LCP_i bases; the rows it ends with are the occurrences that
// reach that maximum, and they are contiguous in suffix order, so the new x0 is the old one plus the number of rows in front
// of the first of them (what kb_extend's `gt` and `primary` terms add up to, step by step).
KB_HD int kb_multi_tail(const KbIndexDev& ix, const KbPk* rd, u64 row, u32 nrows, int done, int cur, int lim, bool* fail, u32* blocks, u32* first, u32* count)
{
	int best = -1; bool bfail = false; u32 bfirst = 0, bcount = 0;
	for (u32 i = 0; i < nrows; i++)
	{
		bool f; const int m = kb_unique_tail(ix, rd, row + i, done, cur, lim, &f, blocks);
		if (m > best) { best = m; bfail = f; bfirst = i; bcount = 1; }
		else if (m == best) bcount++;
	}
	*fail = bfail; *first = bfirst; *count = bcount;
	return best;
}

// One read: all searches of IdentifySeedPairs_FastMode (:49) or _SensitiveMode (:132) with BWT_Search (:140-170) inlined.
// The lanes of a warp each own one read and advance in lock-step, one extension per trip. Everything that is not an
// extension (closing a search: :172-181 and the caller's bookkeeping; opening the next one) is kept out of the extension
// loop and done for several lanes at once: a lane whose search has ended parks until KB_SEED_QUORUM lanes are parked (or
// nobody is searching), then all parked lanes close and reopen together. Per-lane results do not depend on the schedule.
// The read is walked through its packed words (one 16-byte load per 32 bases); a search whose first K bases are clean and
// inside its limit starts from the seeding table instead of K-1 extension steps (identical state by construction).
// Records the searches that will yield seeds (len >= MinSeedLength and interval size <= OCC_Thr 50, bwt_search.cpp:3,172-176).
#define KB_SEED_QUORUM 6
#if defined(__CUDA_ARCH__)
#define KB_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))
#else
#define KB_BALLOT(p) ((p) ? 1u : 0u)
#endif
template <class ROW>
KB_HD void kb_seed_read(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, bool valid, u32* w_steps, u32* w_blocks)
{
	const KbPk* rd = valid ? kb_pk_read(bt, r) : bt.pk;
	const int rlen = valid ? (int)(bt.seq_off[r + 1] - bt.seq_off[r]) : 0;
	KbHit* hits = bt.hits + (size_t)(valid ? r : 0) * bt.max_hits;
	int nh = 0, ns = 0, pos = 0, cur = 0, lim = 0, stop = 30, len = 0;
	const int end = rlen - pm.min_seed, K = (ix.ktab != nullptr && ix.ktab_k <= pm.min_seed) ? ix.ktab_k : 0;
	u32 steps = 0, blocks = 0;
	ROW x0 = 0, x1 = 0, x2 = 0;
	bool searching = false, closing = false, finished = !valid, ovf = false;
	int cw = -1; u64 ccode = 0; u32 cn4 = 0;   // the packed word under the cursor
#if defined(__CUDA_ARCH__)
	const u32 quorum = KB_SEED_QUORUM;
#else
	const u32 quorum = 1;
#endif
	while (KB_BALLOT(!finished))
	{
		// parked lanes: close the search that ended, open the next one
		while (!finished && !searching)
		{
			if (closing)
			{
				closing = false;
				bool hit = len >= pm.min_seed && x2 <= (ROW)50;
				if (hit)
				{
					if (nh < bt.max_hits) { KbHit h; h.x0 = (u64)x0; h.rpos = (u32)pos; h.len_freq = ((u32)len << 8) | (u32)x2; hits[nh++] = h; ns += (int)x2; }
					else ovf = true;
				}
				if (pm.pacbio) { int adv = hit ? len : pm.min_seed; pos += adv; stop += adv; if (stop > rlen) stop = rlen; }
				else pos += len + 1;
			}
			int p = 4;
			while (pos < end)
			{
				if ((pos >> 5) != cw) { cw = pos >> 5; KbPk w = kb_load_pk(rd + cw); ccode = w.code; cn4 = w.n4; }
				const int o = pos & 31;
				p = (int)((ccode >> (62 - 2 * o)) & 3u) | (int)(((cn4 >> (31 - o)) & 1u) << 2);
				if (p <= 3) break;
				pos++; stop++;
			}
			if (pos >= end) { finished = true; break; }
			lim = pm.pacbio ? (stop < rlen ? stop : rlen) : rlen;
			bool seeded = false;
			if (K > 0 && pos + K <= lim)
			{
				const KbPk w = kb_read_win(rd, pos);
				if ((w.n4 >> (32 - K)) == 0)
				{
					const KbKtabE e = kb_load_ktab(ix.ktab + (u32)(w.code >> (64 - 2 * K)), ix.ld_hint); blocks++;
					seeded = true;
					if (e.x2 != 0) { x0 = (ROW)e.x0; x1 = (ROW)e.x1; x2 = (ROW)e.x2; cur = pos + K; steps += (u32)(K - 1); searching = true; }
					else { closing = true; len = (int)e.flen; steps += e.flen; x2 = 0; }
				}
			}
			if (!seeded) { x0 = (ROW)ix.L2[p] + 1; x1 = (ROW)ix.L2[3 - p] + 1; x2 = (ROW)(ix.L2[p + 1] - ix.L2[p]); cur = pos + 1; searching = true; }
		}
		// extension trips until enough lanes are parked
		u32 parked, active;
		do
		{
			if (searching)
			{
				bool ended = true;
				if (cur < lim && x2 == (ROW)1 && ix.sa_full != nullptr)
				{
					bool fail; const int m = kb_unique_tail(ix, rd, (u64)x0, cur - pos, cur, lim, &fail, &blocks);
					cur += m; steps += (u32)m + (fail ? 1u : 0u);
				}
				else if (cur < lim)
				{
					if ((cur >> 5) != cw) { cw = cur >> 5; KbPk w = kb_load_pk(rd + cw); ccode = w.code; cn4 = w.n4; }
					const int o = cur & 31;
					const int c = (int)((ccode >> (62 - 2 * o)) & 3u) | (int)(((cn4 >> (31 - o)) & 1u) << 2);
					if (c <= 3)
					{
						steps++;
						if (kb_extend(ix, x0, x1, x2, c, &blocks)) { cur++; ended = false; }
					}
				}
				if (ended) { searching = false; closing = true; len = cur - pos; }
			}
			active = KB_BALLOT(searching);
			parked = KB_BALLOT(!searching && !finished);
		} while (active != 0 && (u32)KB_POPCLL((u64)parked) < quorum);
	}
	if (!valid) return;
	*w_steps += steps; *w_blocks += blocks;
	if (ovf) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_HITS);
	bt.n_hits[r] = nh; bt.n_seeds[r] = ns;
	u32 off = KB_ALLOC(&bt.counters[0], (u32)ns);
	bt.seed_off[r] = off;
	if ((u64)off + (u64)ns > (u64)bt.cap_segs) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEEDS);
	KB_MAX_U32(&bt.counters[5], (u32)ns);
}

// ---- lane queue -------------------------------------------------------------------------------------------------------
// With kb_unique_tail most reads need a dozen extension trips, a read inside a repeat family still needs ~150: one read per
// lane leaves 31 lanes idle behind the slowest. Here a lane that has finished its read draws the next one from its warp's
// range (Q::next), so the warp lasts as long as its share of the work, not as long as its slowest read. Per outer iteration:
// one pass in which every lane without a running search finishes a one-row search against the text, closes it, and opens the
// next search (table lookup) or finishes the read and draws the next one -- the lanes of a pass run converged; a pass waits
// for `qp` such lanes unless nobody is walking -- and then extension trips for the lanes that walk the index: they wait for
// `qs` walkers while passes can still be filled, and run at least `min_trips` trips so that a long walk is not throttled to
// one step per pass (r15/r16 A/B: qp 8, qs 4, 6 trips on an L2-resident index, 2 on a large one).
// Results per read are identical to kb_seed_read whatever the schedule.
#define KB_SEED_TRIPS 4
// A lane comes back to the packed words of its read for every search (cursor word, table window, the windows of kb_unique_tail): ~50
// 16-byte loads per 150-bp read, which the random index traffic keeps evicting from L1 and L2 -- ncu r32 (C3): the loads that are NOT
// Occ blocks, table or SA entries were four fifths of the L2 read misses of the kernel. A read of up to 32 * (KB_SEED_STAGE - 1)
// bases is therefore copied once into the lane's KB_SEED_STAGE words of shared memory (Q::stage(), streamed past L1) and walked there;
// longer reads (C5) stay where they are. The stride of 7 words = 28 banks keeps the lanes' 16-byte accesses free of bank conflicts.
#define KB_SEED_STAGE 7
struct KbSeedOne   // one read per lane: host emulation, and the reference for the queue
{
	int r; KbPk buf[KB_SEED_STAGE]; bool staged = true;
	KB_HD KbPk* stage() { return staged ? buf : nullptr; }
	KB_HD int next() { int v = r; r = -1; return v; }
	KB_HD void store(const KbBatchDev& bt, int rd, int ns)   // the read's slice of the seed arena, from the grid-wide cursor
	{
		u32 off = KB_ALLOC(&bt.counters[0], (u32)ns);
		bt.seed_off[rd] = off;
		if ((u64)off + (u64)ns > (u64)bt.cap_segs) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEEDS);
		KB_MAX_U32(&bt.counters[5], (u32)ns);
	}
};
template <class ROW, class Q>
KB_HD void kb_seed_lane(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, Q& q, u32* w_steps, u32* w_blocks, int qp = KB_SEED_QUORUM, int qs = 1, int min_trips_arg = KB_SEED_TRIPS, int tail_max = 1)
{
	const ROW tmax = (ROW)(tail_max < 1 ? 1 : tail_max);
	const int K = (ix.ktab != nullptr && ix.ktab_k <= pm.min_seed) ? ix.ktab_k : 0;
	const bool tails = ix.sa_full != nullptr;
	int r = q.next();
	const KbPk* rd = bt.pk; KbHit* hits = bt.hits; int rlen = 0, end = 0;
	int nh = 0, ns = 0, pos = 0, cur = 0, lim = 0, stop = 30, len = 0;
	u32 steps = 0, blocks = 0;
	ROW x0 = 0, x1 = 0, x2 = 0;
	bool searching = false, closing = false, tail = false, finished = r < 0, ovf = false, fresh = !finished;
	int cw = -1; u64 ccode = 0; u32 cn4 = 0;   // the packed word under the cursor
#if defined(__CUDA_ARCH__)
	const u32 quorum_max = (u32)qp, squorum_max = (u32)qs; const int min_trips = min_trips_arg;
#else
	const u32 quorum_max = 1, squorum_max = 1; const int min_trips = 1; (void)qp; (void)qs; (void)min_trips_arg;
#endif
	while (KB_BALLOT(!finished))
	{
		// a pass is worth its ~400 instructions when enough lanes take part (or nobody is walking the index); the same for trips.
		// "Enough" is relative to the lanes that still have a read: a warp with 8 live lanes (long reads: 50 k x 7 kbp are 8 reads per warp;
		// or the tail of any warp's range) must not wait for 8 parked ones -- r26, C5: the seeding kernel lasted 15 ms at one wave.
		const u32 parked0 = KB_BALLOT(!searching && !finished), active0 = KB_BALLOT(searching);
		const u32 live = (u32)KB_POPCLL((u64)(parked0 | active0));
		const u32 quorum = quorum_max < (live + 1) / 2 ? quorum_max : (live + 1) / 2, squorum = squorum_max < (live + 1) / 2 ? squorum_max : (live + 1) / 2;
		const bool do_pass = (u32)KB_POPCLL((u64)parked0) >= quorum || active0 == 0;
		if (do_pass && !finished && !searching)
		{
			if (fresh)   // a new read
			{
				fresh = false;
				rd = kb_pk_read(bt, r); rlen = (int)(bt.seq_off[r + 1] - bt.seq_off[r]); hits = bt.hits + (size_t)r * bt.max_hits;
				KbPk* const st = q.stage();
				if (st != nullptr && rlen <= 32 * (KB_SEED_STAGE - 1))
				{
					const int nw = ((rlen > 0 ? rlen - 1 : 0) >> 5) + 2;   // every word a 32-base window starting inside the read can reach
					for (int k = 0; k < nw; k++) st[k] = kb_load_pk_once(rd + k);
					rd = st;
				}
				end = rlen - pm.min_seed; nh = 0; ns = 0; pos = 0; stop = 30; cw = -1; ovf = false; closing = false; tail = false;
			}
			if (tail)
			{
				tail = false;
				bool fail; int m;
				if (x2 == (ROW)1) m = kb_unique_tail(ix, rd, (u64)x0, cur - pos, cur, lim, &fail, &blocks);
				else { u32 first, count; m = kb_multi_tail(ix, rd, (u64)x0, (u32)x2, cur - pos, cur, lim, &fail, &blocks, &first, &count); x0 += (ROW)first; x2 = (ROW)count; }
				cur += m; steps += (u32)m + (fail ? 1u : 0u); len = cur - pos; closing = true;
			}
			if (closing)
			{
				closing = false;
				bool hit = len >= pm.min_seed && x2 <= (ROW)50;
				if (hit)
				{
					if (nh < bt.max_hits) { KbHit h; h.x0 = (u64)x0; h.rpos = (u32)pos; h.len_freq = ((u32)len << 8) | (u32)x2; hits[nh++] = h; ns += (int)x2; }
					else ovf = true;
				}
				if (pm.pacbio) { int adv = hit ? len : pm.min_seed; pos += adv; stop += adv; if (stop > rlen) stop = rlen; }
				else pos += len + 1;
			}
			int p = 4;
			while (pos < end)
			{
				if ((pos >> 5) != cw) { cw = pos >> 5; KbPk w = kb_load_pk(rd + cw); ccode = w.code; cn4 = w.n4; }
				const int o = pos & 31;
				p = (int)((ccode >> (62 - 2 * o)) & 3u) | (int)(((cn4 >> (31 - o)) & 1u) << 2);
				if (p <= 3) break;
				pos++; stop++;
			}
			if (pos >= end)   // the read is done: its counts and its slice of the seed arena, then the next read of the warp's range
			{
				if (ovf) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_HITS);
				bt.n_hits[r] = nh; bt.n_seeds[r] = ns;
				q.store(bt, r, ns);
				r = q.next();
				if (r < 0) finished = true; else fresh = true;
			}
			else
			{
				lim = pm.pacbio ? (stop < rlen ? stop : rlen) : rlen;
				bool seeded = false;
				if (K > 0 && pos + K <= lim)
				{
					const KbPk w = kb_read_win(rd, pos);
					if ((w.n4 >> (32 - K)) == 0)
					{
						const KbKtabE e = kb_load_ktab(ix.ktab + (u32)(w.code >> (64 - 2 * K)), ix.ld_hint); blocks++;
						seeded = true;
						if (e.x2 != 0) { x0 = (ROW)e.x0; x1 = (ROW)e.x1; x2 = (ROW)e.x2; cur = pos + K; steps += (u32)(K - 1); searching = true; }
						else { closing = true; len = (int)e.flen; steps += e.flen; x2 = 0; }
					}
				}
				if (!seeded) { x0 = (ROW)ix.L2[p] + 1; x1 = (ROW)ix.L2[3 - p] + 1; x2 = (ROW)(ix.L2[p + 1] - ix.L2[p]); cur = pos + 1; searching = true; }
				if (searching && tails && x2 <= tmax && cur < lim) { searching = false; tail = true; }   // straight to the text
			}
		}
		// extension trips
		u32 parked, active; int trips = 0;
		{
			const u32 a1 = KB_BALLOT(searching), p1 = KB_BALLOT(!searching && !finished);
			if (a1 == 0 || (do_pass && (u32)KB_POPCLL((u64)a1) < squorum && (u32)KB_POPCLL((u64)p1) >= quorum)) continue;   // let more lanes join the walk first
		}
		do
		{
			if (searching)
			{
				bool ended = true;
				if (cur < lim)
				{
					if (tails && x2 <= tmax) { tail = true; }
					else
					{
						if ((cur >> 5) != cw) { cw = cur >> 5; KbPk w = kb_load_pk(rd + cw); ccode = w.code; cn4 = w.n4; }
						const int o = cur & 31;
						const int c = (int)((ccode >> (62 - 2 * o)) & 3u) | (int)(((cn4 >> (31 - o)) & 1u) << 2);
						if (c <= 3)
						{
							steps++;
							if (kb_extend(ix, x0, x1, x2, c, &blocks)) { cur++; ended = false; }
						}
					}
				}
				if (ended) { searching = false; if (!tail) { closing = true; len = cur - pos; } }
			}
			trips++;
			active = KB_BALLOT(searching);
			parked = KB_BALLOT(!searching && !finished);
		} while (active != 0 && ((u32)KB_POPCLL((u64)parked) < quorum || trips < min_trips));
	}
	*w_steps += steps; *w_blocks += blocks;
}

// One (read, hit): resolve the SA interval to text positions, in SA-row order (bwt_search.cpp:176-179).
KB_HD void kb_locate_hit(const KbIndexDev& ix, const KbBatchDev& bt, int r, int h, u32* w_lf)
{
	if (h >= bt.n_hits[r]) return;
	const KbHit* hits = bt.hits + (size_t)r * bt.max_hits;
	u32 base = bt.seed_off[r];
	for (int i = 0; i < h; i++) base += hits[i].len_freq & 0xFF;
	KbHit ht = hits[h];
	int freq = (int)(ht.len_freq & 0xFF), len = (int)(ht.len_freq >> 8);
	if ((u64)base + (u64)freq > (u64)bt.cap_segs) return;   // overflow already flagged by kb_seed_read
	for (int i = 0; i < freq; i++)
	{
		u32 steps; u64 g = kb_sa(ix, ht.x0 + (u64)i, &steps); *w_lf += steps;
		KbSeg s; s.gpos = (i64)g; s.rpos = (i32)ht.rpos; s.rlen = len; s.glen = len; s.simple = 1;
		bt.segs[base + i] = s;
	}
}

#endif
