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

#if defined(__CUDA_ARCH__)
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

// Occ(k, .) for all four bases: number of each base in BWT rows [0..k] with '$' skipped. k != ~0.   (bwt_occ4 :68)
KB_HD void kb_occ4(const KbIndexDev& ix, u64 k, u64 out[4])
{
	u64 r = k - (k >= ix.primary);
	const uint4* p = reinterpret_cast<const uint4*>(ix.occ) + ((r >> 6) << 1);
	uint4 c = KB_LDG4(p), w = KB_LDG4(p + 1);
	int n = (int)(r & 63) + 1;
	u32 cnt[4] = {c.x, c.y, c.z, c.w};
	kb_count32(((u64)w.x << 32) | w.y, n < 32 ? n : 32, cnt);
	if (n > 32) kb_count32(((u64)w.z << 32) | w.w, n - 32, cnt);
	out[0] = cnt[0]; out[1] = cnt[1]; out[2] = cnt[2]; out[3] = cnt[3];
}

// ---- one 32-byte Occ block = one DRAM sector, fetched with a single 256-bit load (LDG.E.256 on sm_100a) ----
struct KbBlk { u32 c0, c1, c2, c3, w0, w1, w2, w3; };
KB_HD KbBlk kb_load_blk(const uint32_t* occ, u64 blk)
{
	KbBlk b; const uint32_t* p = occ + (blk << 3);
#if defined(__CUDA_ARCH__)
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(b.c0), "=r"(b.c1), "=r"(b.c2), "=r"(b.c3), "=r"(b.w0), "=r"(b.w1), "=r"(b.w2), "=r"(b.w3) : "l"(p));
#else
	b.c0 = p[0]; b.c1 = p[1]; b.c2 = p[2]; b.c3 = p[3]; b.w0 = p[4]; b.w1 = p[5]; b.w2 = p[6]; b.w3 = p[7];
#endif
	return b;
}

// For base b: eq = Occ(row, b), gt = sum over bases j > b of Occ(row, j), where `off` is the row's offset in the block.
// Two popcounts per 32 symbols instead of one per base: the interval update only needs these two sums.
KB_HD void kb_rank_eq_gt(const KbBlk& k, int off, int b, u32* eq, u32* gt)
{
	// all selections by `b` are mask arithmetic (see kb_nt4 for why): BL/BH broadcast b's low/high bit to every symbol slot
	const u64 M5 = 0x5555555555555555ull;
	u32 n = (u32)off + 1u, n0 = n < 32u ? n : 32u, n1 = n - n0;
	u64 pm0 = M5 & ~((~0ull >> n0) >> n0);            // first n0 symbols of word 0 (n0 = 1..32)
	u64 pm1 = M5 & ~((~0ull >> n1) >> n1);            // first n1 symbols of word 1 (n1 = 0..32)
	u64 BL = M5 & (0ull - (u64)((u32)b & 1u)), BH = M5 & (0ull - (u64)(((u32)b >> 1) & 1u));
	u64 W0 = ((u64)k.w0 << 32) | k.w1, W1 = ((u64)k.w2 << 32) | k.w3;
	u64 lo0 = W0 & M5, hi0 = (W0 >> 1) & M5, lo1 = W1 & M5, hi1 = (W1 >> 1) & M5;
	u64 e0 = ~((lo0 ^ BL) | (hi0 ^ BH)) & pm0, e1 = ~((lo1 ^ BL) | (hi1 ^ BH)) & pm1;                     // symbol == b
	u64 g0 = ((hi0 & ~BH) | (~(hi0 ^ BH) & lo0 & ~BL)) & pm0, g1 = ((hi1 & ~BH) | (~(hi1 ^ BH) & lo1 & ~BL)) & pm1;   // symbol > b
	u32 m0 = 0u - (u32)(b == 0), m1 = 0u - (u32)(b == 1), m2 = 0u - (u32)(b == 2), m3 = 0u - (u32)(b == 3);
	u32 ceq = (k.c0 & m0) | (k.c1 & m1) | (k.c2 & m2) | (k.c3 & m3);
	u32 cgt = ((k.c1 + k.c2 + k.c3) & m0) | ((k.c2 + k.c3) & m1) | (k.c3 & m2);
	*eq = ceq + (u32)KB_POPCLL(e0) + (u32)KB_POPCLL(e1);
	*gt = cgt + (u32)KB_POPCLL(g0) + (u32)KB_POPCLL(g1);
}

// ---- one forward extension of a bi-interval by base c (BWT_Search :151-166 with bwt_2occ4 :87) -------------------------
// b = 3 - c is the base looked up in the BWT. Returns false (state untouched) when the extended interval is empty.
// Rows k' and l' usually share one 32-byte block (always, once the interval is narrow): then the block is loaded once,
// Occ(k,b) comes from a prefix mask and the two differences Occ(l,.)-Occ(k,.) from a range mask.
KB_HD bool kb_extend(const KbIndexDev& ix, u64& x0, u64& x1, u64& x2, int c, u32* blocks)
{
	const u64 M5 = 0x5555555555555555ull;
	const u64 primary = ix.primary;
	const int b = 3 - c;
	const u64 k = x1 - 1, l = k + x2;
	const u64 rk = k - (k >= primary), rl = l - (l >= primary);
	u32 ek, n2, gt;
	if ((rk >> 6) == (rl >> 6))
	{
		const KbBlk B = kb_load_blk(ix.occ, rk >> 6); *blocks += 1;
		const u32 nk = (u32)(rk & 63) + 1u, nl = (u32)(rl & 63) + 1u;
		const u32 nk0 = nk < 32u ? nk : 32u, nk1 = nk - nk0, nl0 = nl < 32u ? nl : 32u, nl1 = nl - nl0;
		const u64 pk0 = M5 & ~((~0ull >> nk0) >> nk0), pk1 = M5 & ~((~0ull >> nk1) >> nk1);
		const u64 r0 = M5 & ~((~0ull >> nl0) >> nl0) & ~pk0, r1 = M5 & ~((~0ull >> nl1) >> nl1) & ~pk1;   // rows (k', l']
		const u64 BL = M5 & (0ull - (u64)((u32)b & 1u)), BH = M5 & (0ull - (u64)(((u32)b >> 1) & 1u));
		const u64 W0 = ((u64)B.w0 << 32) | B.w1, W1 = ((u64)B.w2 << 32) | B.w3;
		const u64 lo0 = W0 & M5, hi0 = (W0 >> 1) & M5, lo1 = W1 & M5, hi1 = (W1 >> 1) & M5;
		const u64 e0 = ~((lo0 ^ BL) | (hi0 ^ BH)), e1 = ~((lo1 ^ BL) | (hi1 ^ BH));   // symbol == b (valid under the M5-based masks)
		n2 = (u32)KB_POPCLL(e0 & r0) + (u32)KB_POPCLL(e1 & r1);
		if (n2 == 0) return false;
		const u32 m0 = 0u - (u32)(b == 0), m1 = 0u - (u32)(b == 1), m2 = 0u - (u32)(b == 2), m3 = 0u - (u32)(b == 3);
		ek = ((B.c0 & m0) | (B.c1 & m1) | (B.c2 & m2) | (B.c3 & m3)) + (u32)KB_POPCLL(e0 & pk0) + (u32)KB_POPCLL(e1 & pk1);
		gt = 0;
		if (x2 > 1)   // a one-row interval that survives holds b itself: nothing greater
		{
			const u64 g0 = (hi0 & ~BH) | (~(hi0 ^ BH) & lo0 & ~BL), g1 = (hi1 & ~BH) | (~(hi1 ^ BH) & lo1 & ~BL);   // symbol > b
			gt = (u32)KB_POPCLL(g0 & r0) + (u32)KB_POPCLL(g1 & r1);
		}
	}
	else
	{
		const KbBlk bk = kb_load_blk(ix.occ, rk >> 6), bl = kb_load_blk(ix.occ, rl >> 6); *blocks += 2;
		u32 gk, el, gl;
		kb_rank_eq_gt(bk, (int)(rk & 63), b, &ek, &gk);
		kb_rank_eq_gt(bl, (int)(rl & 63), b, &el, &gl);
		n2 = el - ek; gt = gl - gk;
		if (n2 == 0) return false;
	}
	x0 = x0 + ((x1 <= primary && x1 + x2 - 1 >= primary) ? 1 : 0) + (u64)gt;
	x1 = ix.L2[b] + 1 + ek; x2 = n2;
	return true;
}

// one entry of the seeding table: BWT_Search over the K bases spelled by `kmer` (first base most significant)
KB_HD KbKtab kb_ktab_entry(const KbIndexDev& ix, u32 kmer, int K)
{
	KbKtab e; e.pad = 0; e.flen = 0;
	int p = (int)((kmer >> (2 * (K - 1))) & 3u);
	u64 x0 = ix.L2[p] + 1, x1 = ix.L2[3 - p] + 1, x2 = ix.L2[p + 1] - ix.L2[p];
	u32 blocks = 0;
	for (int i = 1; i < K; i++)
	{
		int c = (int)((kmer >> (2 * (K - 1 - i))) & 3u);
		if (!kb_extend(ix, x0, x1, x2, c, &blocks)) { e.x0 = 0; e.x1 = 0; e.x2 = 0; e.flen = (u32)i; return e; }
	}
	e.x0 = x0; e.x1 = x1; e.x2 = (u32)x2;
	if (x2 == 0) { e.x0 = 0; e.x1 = 0; e.flen = (u32)K; }   // a base that does not occur (K == 1 only)
	return e;
}
KB_HD KbKtab kb_load_ktab(const KbKtab* p)
{
#if defined(__CUDA_ARCH__)
	u32 r[8];
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
	KbKtab e; e.x0 = ((u64)r[1] << 32) | r[0]; e.x1 = ((u64)r[3] << 32) | r[2]; e.x2 = r[4]; e.flen = r[5]; e.pad = 0; return e;
#else
	return *p;
#endif
}

// LF step (bwt_invPsi :120 with bwt_occ :44 folded in; one 32-byte block per step)
KB_HD u64 kb_lf(const KbIndexDev& ix, u64 k)
{
	if (k == ix.primary) return 0;
	u64 r = k - (k > ix.primary);
	KbBlk bk = kb_load_blk(ix.occ, r >> 6);
	int off = (int)(r & 63);
	u32 word = off < 16 ? bk.w0 : (off < 32 ? bk.w1 : (off < 48 ? bk.w2 : bk.w3));
	int sym = (word >> ((~off & 15) << 1)) & 3;
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

// One read: all searches of IdentifySeedPairs_FastMode (:49) or _SensitiveMode (:132) with BWT_Search (:140-170) inlined as a
// FLAT state machine: every trip of the single loop performs (at most) one extension step, so the lanes of a warp stay in
// lock-step across search boundaries instead of waiting for the longest search of the warp (nested loops cost 4-5x here).
// The read is walked through its packed words (one 16-byte load per 32 bases); a search whose first K bases are clean and
// inside its limit starts from the seeding table instead of K-1 extension steps (identical state by construction).
// Records the searches that will yield seeds (len >= MinSeedLength and interval size <= OCC_Thr 50, bwt_search.cpp:3,172-176).
KB_HD void kb_seed_read(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, u32* w_steps, u32* w_blocks)
{
	const KbPk* rd = kb_pk_read(bt, r);
	const int rlen = (int)(bt.seq_off[r + 1] - bt.seq_off[r]);
	KbHit* hits = bt.hits + (size_t)r * bt.max_hits;
	int nh = 0, ns = 0, pos = 0, cur = 0, lim = 0, stop = 30;
	const int end = rlen - pm.min_seed, K = (ix.ktab != nullptr && ix.ktab_k <= pm.min_seed) ? ix.ktab_k : 0;
	u32 steps = 0, blocks = 0;
	u64 x0 = 0, x1 = 0, x2 = 0;
	bool searching = false, ovf = false;
	int cw = -1; u64 ccode = 0; u32 cn4 = 0;   // the packed word under the cursor
	while (searching || pos < end)
	{
		bool ended = false; int len = 0;
		if (!searching)
		{
			if ((pos >> 5) != cw) { cw = pos >> 5; KbPk w = kb_load_pk(rd + cw); ccode = w.code; cn4 = w.n4; }
			const int o = pos & 31;
			const int p = (int)((ccode >> (62 - 2 * o)) & 3u) | (int)(((cn4 >> (31 - o)) & 1u) << 2);
			if (p > 3) { pos++; stop++; continue; }
			lim = pm.pacbio ? (stop < rlen ? stop : rlen) : rlen;
			searching = true;
			bool seeded = false;
			if (K > 0 && pos + K <= lim)
			{
				const KbPk w = kb_read_win(rd, pos);
				if ((w.n4 >> (32 - K)) == 0)
				{
					const KbKtab e = kb_load_ktab(ix.ktab + (u32)(w.code >> (64 - 2 * K))); blocks++;
					seeded = true;
					if (e.x2 != 0) { x0 = e.x0; x1 = e.x1; x2 = e.x2; cur = pos + K; steps += (u32)(K - 1); }
					else { ended = true; len = (int)e.flen; steps += e.flen; x2 = 0; }
				}
			}
			if (!seeded) { x0 = ix.L2[p] + 1; x1 = ix.L2[3 - p] + 1; x2 = ix.L2[p + 1] - ix.L2[p]; cur = pos + 1; }
		}
		if (!ended)
		{
			ended = true;
			if (cur < lim)
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
			len = cur - pos;
		}
		if (ended)
		{
			bool hit = len >= pm.min_seed && (int)x2 <= 50;
			if (hit)
			{
				if (nh < bt.max_hits) { KbHit h; h.x0 = x0; h.rpos = (u32)pos; h.len_freq = ((u32)len << 8) | (u32)x2; hits[nh++] = h; ns += (int)x2; }
				else ovf = true;
			}
			if (pm.pacbio) { int adv = hit ? len : pm.min_seed; pos += adv; stop += adv; if (stop > rlen) stop = rlen; }
			else pos += len + 1;
			searching = false;
		}
	}
	*w_steps += steps; *w_blocks += blocks;
	if (ovf) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_HITS);
	bt.n_hits[r] = nh; bt.n_seeds[r] = ns;
	u32 off = KB_ATOMIC_ADD(&bt.counters[0], (u32)ns);
	bt.seed_off[r] = off;
	if ((u64)off + (u64)ns > (u64)bt.cap_segs) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEEDS);
	KB_ATOMIC_MAX(&bt.counters[5], (u32)ns);
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
