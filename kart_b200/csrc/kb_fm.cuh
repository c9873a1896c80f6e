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
#define KB_ATOMIC_MAX(p, v) kb_host_max((p), (v))
#endif

// nst_nt4_table (src/BWT_Index/bntseq.c:40): A/a 0, C/c 1, G/g 2, T/t 3, everything else 4
KB_HD int kb_nt4(u8 c)
{
	c &= 0xDF;   // fold case: 'a'(0x61)->'A'(0x41) ; other bytes may alias but only onto non-ACGT codes or the same letters
	return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : 4)));
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

// LF step (bwt_invPsi :120 with bwt_occ :44 folded in; one 32-byte block per step)
KB_HD u64 kb_lf(const KbIndexDev& ix, u64 k)
{
	if (k == ix.primary) return 0;
	u64 r = k - (k > ix.primary);
	const uint4* p = reinterpret_cast<const uint4*>(ix.occ) + ((r >> 6) << 1);
	uint4 c = KB_LDG4(p), w = KB_LDG4(p + 1);
	int off = (int)(r & 63);
	u32 word = off < 16 ? w.x : (off < 32 ? w.y : (off < 48 ? w.z : w.w));
	int sym = (word >> ((~off & 15) << 1)) & 3;
	u32 cnt[4] = {c.x, c.y, c.z, c.w};
	int n = off + 1;
	kb_count32(((u64)w.x << 32) | w.y, n < 32 ? n : 32, cnt);
	if (n > 32) kb_count32(((u64)w.z << 32) | w.w, n - 32, cnt);
	return ix.L2[sym] + cnt[sym];
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

struct KbSearch { u64 x0, x2; int len; u32 steps, blocks; };

// Forward extension of a bi-interval from seq[start] while the match is non-empty (BWT_Search :140-170).
KB_HD KbSearch kb_search(const KbIndexDev& ix, const u8* seq, int start, int stop)
{
	KbSearch o; o.steps = 0; o.blocks = 0;
	int p = kb_nt4(seq[start]), pos;
	u64 x0 = ix.L2[p] + 1, x1 = ix.L2[3 - p] + 1, x2 = ix.L2[p + 1] - ix.L2[p];
	for (pos = start + 1; pos < stop; pos++)
	{
		int c = kb_nt4(seq[pos]);
		if (c > 3) break;
		u64 tk[4], tl[4], k = x1 - 1, l = x1 - 1 + x2;
		kb_occ4(ix, k, tk); kb_occ4(ix, l, tl);
		o.steps++; o.blocks += 1 + (((k - (k >= ix.primary)) >> 6) != ((l - (l >= ix.primary)) >> 6));
		int b = 3 - c;
		u64 n2 = tl[b] - tk[b];
		if (n2 == 0) break;
		u64 n0 = x0 + ((x1 <= ix.primary && x1 + x2 - 1 >= ix.primary) ? 1 : 0);
		for (int j = 3; j > b; j--) n0 += tl[j] - tk[j];
		x0 = n0; x1 = ix.L2[b] + 1 + tk[b]; x2 = n2;
	}
	o.x0 = x0; o.x2 = x2; o.len = pos - start;
	return o;
}

// One read: all searches of IdentifySeedPairs_FastMode (:49) or _SensitiveMode (:132); records the searches that
// will yield seeds (len >= MinSeedLength and interval size <= OCC_Thr 50, bwt_search.cpp:3,172-176).
KB_HD void kb_seed_read(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, u32* w_steps, u32* w_blocks)
{
	const u8* seq = bt.seq + bt.seq_off[r];
	int rlen = (int)(bt.seq_off[r + 1] - bt.seq_off[r]);
	KbHit* hits = bt.hits + (size_t)r * bt.max_hits;
	int nh = 0, ns = 0, pos = 0, end = rlen - pm.min_seed, stop = 30;
	bool ovf = false;
	while (pos < end)
	{
		if (kb_nt4(seq[pos]) > 3) { pos++; stop++; continue; }
		KbSearch s = kb_search(ix, seq, pos, pm.pacbio ? stop : rlen);
		*w_steps += s.steps; *w_blocks += s.blocks;
		bool hit = s.len >= pm.min_seed && (int)s.x2 <= 50;
		if (hit)
		{
			if (nh < bt.max_hits) { KbHit h; h.x0 = s.x0; h.rpos = (u32)pos; h.len_freq = ((u32)s.len << 8) | (u32)s.x2; hits[nh++] = h; ns += (int)s.x2; }
			else ovf = true;
		}
		if (pm.pacbio)
		{
			int adv = hit ? s.len : pm.min_seed;
			pos += adv; stop += adv; if (stop > rlen) stop = rlen;
		}
		else pos += s.len + 1;
	}
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
