// Gap filling, CIGAR construction and per-read reports (one thread per read, private scratch arena in HBM).
// Replaces: IdentifyNormalPairs src/AlignmentCandidates.cpp:420 (+ :235,:273,:323,:382), GenerateSimplePairsFromFragmentPair
// src/KmerAnalysis.cpp:164 (+ :56,:104,:132), nw_alignment src/nw_alignment.cpp:18, GenerateNormalPairAlignment src/tools.cpp:142,
// Process{Normal,Head,Tail}SequencePair :225,:292,:344, AddNewCigarElements :49, CheckLocalAlignmentQuality :255,
// GenMappingReport src/AlignmentCandidates.cpp:624, CheckCoordinateValidity :582, GenCoordinateInfo :515, GenerateCIGAR :492.
//
// Alignments are never materialised as gapped strings: every consumer in the reference (CIGAR run-length encoding,
// identity count, the head/tail quality test, leading/trailing gap stripping) is a function of the run list
// (type,len) plus the number of identical aligned characters, so that is what is carried around.
#ifndef KB_ALIGN_CUH
#define KB_ALIGN_CUH
#include "kb_cand.cuh"

// ---- per-thread stack arena --------------------------------------------------------------------
struct KbArena
{
	u8* base; u64 used, cap; bool ovf;
	KB_HD void* alloc(u64 bytes)
	{
		bytes = (bytes + 15) & ~(u64)15;
		if (used + bytes > cap) { ovf = true; return nullptr; }
		void* p = base + used; used += bytes; return p;
	}
};

// ---- reference text: Kart's 2G text = forward strand + reverse complement (src/bwt_index.cpp:194-213) ----
KB_HD int kb_ref_code(const KbIndexDev& ix, i64 p)   // 0..3, or 4 outside [0,2G) (undefined behaviour in the reference)
{
	if (p < 0 || p >= ix.G2) return 4;
	bool rev = p >= ix.G; i64 f = rev ? ix.G2 - 1 - p : p;
	int c = (KB_LDG(ix.pac + (f >> 2)) >> ((~f & 3) << 1)) & 3;
	return rev ? 3 - c : c;
}
KB_HD u8 kb_code_char(int c) { return c == 0 ? 'A' : (c == 1 ? 'C' : (c == 2 ? 'G' : (c == 3 ? 'T' : 'N'))); }
// 32 characters of the 2G text starting at p as 2-bit codes, MSB first; *inval flags (bit 31-i) the positions outside [0,2G)
KB_HD u64 kb_ref_win(const KbIndexDev& ix, i64 p, u32* inval)
{
	const u64 M5 = 0x5555555555555555ull;
	*inval = 0;
	if (p >= 0 && p + 32 <= ix.G)
	{
		const u64* w = ix.ref64 + (p >> 5); const int s = (int)(p & 31) * 2;
		u64 hi = KB_LDG(w); if (s == 0) return hi;
		return (hi << s) | (KB_LDG(w + 1) >> (64 - s));
	}
	if (p >= ix.G && p + 32 <= ix.G2)
	{
		const i64 f = ix.G2 - 32 - p;   // forward bases f..f+31 are the window read backwards
		const u64* w = ix.ref64 + (f >> 5); const int s = (int)(f & 31) * 2;
		u64 v = KB_LDG(w); if (s) v = (v << s) | (KB_LDG(w + 1) >> (64 - s));
#if defined(__CUDA_ARCH__)
		v = __brevll(v);
#else
		{ u64 x = v; x = ((x >> 1) & M5) | ((x & M5) << 1); x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
		  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4); v = __builtin_bswap64(x); }
#endif
		v = ((v >> 1) & M5) | ((v & M5) << 1);   // bit reversal swapped the two bits of every base
		return ~v;
	}
	u64 v = 0; u32 bad = 0;
	for (int i = 0; i < 32; i++) { int c = kb_ref_code(ix, p + i); if (c > 3) bad |= 1u << (31 - i); else v |= (u64)c << (62 - 2 * i); }
	*inval = bad;
	return v;
}

// ---- run-list accumulator ----------------------------------------------------------------------
enum { KB_RUN_D = 0, KB_RUN_I = 1, KB_RUN_M = 2 };   // gap in read / gap in genome / aligned column  (CheckLocalAlignmentQuality types 0,1,2)
struct KbRuns
{
	u32* r; int n, cap; int ident, aligned; bool ovf;
	u32 tail;            // the open run (not yet stored); 0 = none. Keeps the read-modify-write of the last run out of memory.
	KB_HD void reset(u32* dst, int capacity) { r = dst; cap = capacity; n = 0; ident = 0; aligned = 0; ovf = false; tail = 0; }
	KB_HD void push(int type, int len)
	{
		if (len <= 0) return;
		if (tail != 0 && (int)(tail & 3) == type) { tail += (u32)len << 2; return; }
		flush();
		tail = ((u32)len << 2) | (u32)type;
	}
	KB_HD void flush()
	{
		if (tail == 0) return;
		if (n >= cap) ovf = true; else r[n++] = tail;
		tail = 0;
	}
};

// ---- cigar list of one candidate (elements exactly as the reference pushes them into cigar_vec) ----
struct KbCigar
{
	u32* e; int n, cap; bool ovf;
	KB_HD void push(int len, int op) { if (n >= cap) { ovf = true; return; } e[n++] = ((u32)len << 4) | (u32)op; }
};

// ================================================================================================
// IdentifyNormalPairs
// ================================================================================================
KB_HD int kb_drop_empty(KbSeg* v, int n) { int w = 0; for (int i = 0; i < n; i++) if (v[i].rlen != 0) v[w++] = v[i]; return w; }

// order[] = indices of v sorted by rpos (insertion sort; lists are short)
KB_HD void kb_order_by_rpos(const KbSeg* v, int n, i32* order)
{
	for (int i = 0; i < n; i++)
	{
		int j = i, key = v[i].rpos;
		while (j > 0 && v[order[j - 1]].rpos > key) { order[j] = order[j - 1]; j--; }
		order[j] = i;
	}
}

KB_HD int kb_drop_shared_rpos(KbSeg* v, int n, i32* order)   // RemoveTandemRepeatSeeds :235
{
	if (n < 2) return n;
	kb_order_by_rpos(v, n, order);
	bool any = false;
	for (int i = 0; i < n;)
	{
		int j = i + 1; while (j < n && v[order[j]].rpos == v[order[i]].rpos) j++;
		if (j - i > 1) { any = true; for (int k = i; k < j; k++) v[order[k]].rlen = v[order[k]].glen = 0; }
		i = j;
	}
	return any ? kb_drop_empty(v, n) : n;
}

KB_HD int kb_drop_translocated(KbSeg* v, int n, i32* order)   // RemoveTranslocatedSeeds :273
{
	if (n < 2) return n;
	kb_order_by_rpos(v, n, order);
	bool any = false;
	for (int i = 0; i < n; i++)
	{
		if (v[order[i]].rpos == v[i].rpos) continue;
		any = true;
		int hi = order[i];
		for (int j = i + 1; j <= hi; j++) if (order[j] > hi) hi = order[j];
		int s1 = 0, s2 = 0;
		for (int k = i; k <= hi; k++) { if (k < order[k]) s1 += v[order[k]].rlen; else s2 += v[order[k]].rlen; }
		for (int k = i; k <= hi; k++)
		{
			bool kill = (s1 > s2) ? (k > order[k]) : (k < order[k]);
			if (kill) v[order[k]].rlen = v[order[k]].glen = 0;
		}
		i = hi;
	}
	return any ? kb_drop_empty(v, n) : n;
}

KB_HD bool kb_resolve_overlap(KbSeg& a, KbSeg& b)   // CheckSeedOverlapping :323
{
	bool master = true; int ov;
	if ((ov = a.rpos + a.rlen - b.rpos) > 0)
	{
		if (a.rlen < b.rlen) { master = false; if (a.rlen > ov) a.glen = (a.rlen -= ov); else a.rlen = a.glen = 0; }
		else if (b.rlen > ov) { b.rpos += ov; b.gpos += ov; b.glen = (b.rlen -= ov); }
		else b.rlen = b.glen = 0;
	}
	if (a.rlen > 0 && b.rlen > 0 && (ov = (int)(a.gpos + a.glen - b.gpos)) > 0)
	{
		if (a.glen < b.glen) { master = false; if (a.rlen > ov) a.glen = (a.rlen -= ov); else a.rlen = a.glen = 0; }
		else if (b.rlen > ov) { b.rpos += ov; b.gpos += ov; b.glen = (b.rlen -= ov); }
		else b.rlen = b.glen = 0;
	}
	return master;
}

KB_HD int kb_trim_overlaps(KbSeg* v, int n)   // CheckOverlappingSeeds :382
{
	if (n < 2) return n;
	bool any = false;
	for (int i = 0; i < n;)
	{
		if (v[i].rlen > 0)
		{
			int rEnd = v[i].rpos + v[i].rlen - 1; i64 gEnd = v[i].gpos + v[i].glen - 1;
			for (int j = i + 1; j < n; j++)
			{
				if (v[j].rlen == 0) continue;
				if (rEnd < v[j].rpos && gEnd < v[j].gpos) break;
				if (!kb_resolve_overlap(v[i], v[j])) break;
			}
			if (v[i].rlen == 0)
			{
				any = true;
				int p = i - 1; while (p > 0 && v[p].rlen == 0) p--;
				i = p < 0 ? 0 : p;
			}
			else i++;
		}
		else { any = true; i++; }
	}
	return any ? kb_drop_empty(v, n) : n;
}

// Expands the simple pairs in `in` (n entries, (gPos,rPos)-sorted) into the full segment list written to `out`
// (capacity >= 2n+2). glen < 0: whole read against the genome (head/tail get gLen = rLen). Returns the count.
KB_HD int kb_fill_pairs(int rlen, int glen, KbSeg* in, int n, KbSeg* out, i32* order)
{
	int total = 0;
	if (n > 1)
	{
		n = kb_drop_shared_rpos(in, n, order);
		n = kb_drop_translocated(in, n, order);
		n = kb_trim_overlaps(in, n);
		// gap ("normal") pairs in seed order, parked behind the merge output (write index si+gi never reaches read index n+1+gi)
		KbSeg* gaps = out + n + 1; int ng = 0;
		for (int i = 0, j = 1; j < n; i++, j++)
		{
			int rg = in[j].rpos - (in[i].rpos + in[i].rlen); if (rg < 0) rg = 0;
			int gg = (int)(in[j].gpos - (in[i].gpos + in[i].glen)); if (gg < 0) gg = 0;
			if (rg > 0 || gg > 0)
			{
				KbSeg g; g.simple = 0; g.rpos = in[i].rpos + in[i].rlen; g.gpos = in[i].gpos + in[i].glen; g.rlen = rg; g.glen = gg;
				gaps[ng++] = g;
			}
		}
		// std::inplace_merge(:449) == stable merge of the two runs by (gPos,rPos)
		int si = 0, gi = 0;
		while (si < n && gi < ng) { if (kb_less_gpos(gaps[gi], in[si])) out[total++] = gaps[gi++]; else out[total++] = in[si++]; }
		while (si < n) out[total++] = in[si++];
		while (gi < ng) out[total++] = gaps[gi++];
	}
	else { for (int i = 0; i < n; i++) out[i] = in[i]; total = n; }
	if (total > 0)
	{
		int rg = out[0].rpos > 0 ? out[0].rpos : 0;
		int gg = glen > 0 ? (int)out[0].gpos : rg;
		if (rg > 0 || gg > 0)
		{
			KbSeg h; h.simple = 0; h.rpos = 0; h.gpos = out[0].gpos - gg; if (h.gpos < 0) h.gpos = 0;
			h.rlen = rg; h.glen = gg;
			for (int i = total; i > 0; i--) out[i] = out[i - 1];
			out[0] = h; total++;
		}
		const KbSeg t = out[total - 1];
		rg = rlen - (t.rpos + t.rlen);
		gg = glen > 0 ? glen - (int)(t.gpos + t.glen) : rg;
		if (rg > 0 || gg > 0)
		{
			KbSeg e; e.simple = 0; e.rpos = t.rpos + t.rlen; e.gpos = t.gpos + t.glen; e.rlen = rg; e.glen = gg;
			out[total++] = e;
		}
	}
	return total;
}

// ================================================================================================
// 8-mer partition of a fragment pair
// ================================================================================================
#define KB_NOKMER 0xFFFFFFFFu

// w[p] = word id CreateKmerVecFromReadSeq (KmerAnalysis.cpp:56-98) records for position p, KB_NOKMER where it records none.
// This is a literal transcription of the reference's scan, because its bookkeeping after a literal 'N' is off by one
// (the loop increment at :74 runs once more after the restart at :93, so the character right behind the first clean
// window is skipped and every later 8-mer is labelled one position early); results must match that, not the intent.
KB_HD u32 kb_fresh_kmer(const u8* s, u32 pos) { u32 id = 0; for (u32 i = pos; i < pos + 8; i++) id = (id << 2) + (u32)kb_nt4(s[i]); return id; }
KB_HD void kb_kmer_ids(int len, const u8* s, u32* w)
{
	for (int i = 0; i < len; i++) w[i] = KB_NOKMER;
	u32 count = 0, head, tail = 0, n = (u32)(len < 0 ? 0 : len);
	while (count < 8 && tail < n) { if (s[tail++] != 'N') count++; else count = 0; }
	if (count != 8) return;
	head = tail - 8;
	u32 wid = kb_fresh_kmer(s, head); w[head] = wid;
	for (head += 1; tail < n; head++, tail++)
	{
		if (s[tail] != 'N') { wid = ((wid & 0x3FFF) << 2) + (u32)kb_nt4(s[tail]); w[head] = wid; }
		else
		{
			count = 0; tail++;
			while (count < 8 && tail < n) { if (s[tail++] != 'N') count++; else count = 0; }
			if (count != 8) break;
			head = tail - 8; wid = kb_fresh_kmer(s, head); w[head] = wid;
		}
	}
}

// Exact-match runs of common 8-mers, emitted in (PosDiff,rPos) order like GenerateSimplePairsFromCommonKmers(:132):
// a pair (r,g) exists iff both 8-mers exist, ids are equal and |g-r| < max_shift (IdentifyCommonKmers :118).
KB_HD int kb_kmer_pairs(const u32* w1, int len1, const u32* w2, int len2, int max_shift, int min_len, KbSeg* out, int cap, bool* ovf)
{
	int n = 0;
	int n1 = len1 - 7, n2 = len2 - 7;   // number of candidate start positions
	if (n1 <= 0 || n2 <= 0) return 0;
	int dlo = -(n1 - 1), dhi = n2 - 1;
	if (dlo < -(max_shift - 1)) dlo = -(max_shift - 1);
	if (dhi > max_shift - 1) dhi = max_shift - 1;
	for (int d = dlo; d <= dhi; d++)
	{
		int r0 = d < 0 ? -d : 0, r1 = n1 < n2 - d ? n1 : n2 - d;   // r in [r0, r1)
		int run = 0;
		for (int r = r0; r <= r1; r++)
		{
			bool m = r < r1 && w1[r] != KB_NOKMER && w1[r] == w2[r + d];
			if (m) run++;
			else if (run > 0)
			{
				int l = 8 + run - 1;
				if (l >= min_len)
				{
					if (n < cap) { KbSeg s; s.simple = 1; s.rpos = r - run; s.gpos = (i64)(r - run + d); s.rlen = s.glen = l; out[n++] = s; }
					else *ovf = true;
				}
				run = 0;
			}
		}
	}
	return n;
}

// ================================================================================================
// Needleman-Wunsch, integer recurrence with all scores doubled (exactly equivalent to the float DP of
// nw_alignment.cpp because every value there is a multiple of 0.5), 2-bit traceback.
//
// Warp-per-fragment, anti-diagonal wavefront: rows are processed in strips of 32, lane t owns row i0+t+1 and at
// step d computes column j = d-t+1, so all lanes of a strip advance along one anti-diagonal per step. A lane needs
//   left  S[i][j-1], R[i][j-1]   its own previous step (registers)
//   up    S[i-1][j], T[i-1][j]   lane t-1's previous step (exchanged through a double-buffered shared array),
//                                 or the stored last row of the previous strip for lane 0
//   diag  S[i-1][j-1]            the `up` value it saw one step earlier (register)
// The step is written as a function of (warp state, lane state, step, lane) so that the host-emulation build can replay
// a warp by looping over lanes; on the GPU the 32 calls are the 32 lanes and `__syncwarp()` separates the steps.
// ================================================================================================
#define KB_NW_NEG (-131072)
struct KbNwWarp
{
	const u8* c1; const u8* c2; int m, n;
	int* bS[2]; int* bT[2];     // stored boundary row (row i0), only when m > 32: read buffer = cur, written buffer = cur ^ 1
	u8* code2;                  // nt4 codes of c2, 1-based
	u8* tb; u32 stride;         // traceback: 2 bits per cell (bit 0: S==R, bit 1: S==T), 4 cells per byte, row-major
	int xs[2][32], xt[2][32];   // neighbour exchange, indexed [step & 1][lane]
	int i0, h, cur;
	u64 mark, fmark;
};
// per-lane registers: DP state of the lane's row plus the strip constants it needs every step
struct KbNwLane { int left_s, left_r, diag, a, n, h, last; u32 pack; const u8* code2; u8* tbrow; const int* rS; const int* rT; int* wS; int* wT; };

// two-level allocation: the warp's shared-memory pool first, the (L2-latency) HBM arena when the problem is too large
KB_HD void* kb_alloc2(KbArena& fast, KbArena& slow, u64 bytes)
{
	u64 need = (bytes + 15) & ~(u64)15;
	if (fast.used + need <= fast.cap) { void* p = fast.base + fast.used; fast.used += need; return p; }
	return slow.alloc(bytes);
}

// lane 0: storage for one (m x n) problem
KB_HD bool kb_nww_setup(KbNwWarp& w, KbArena& fast, KbArena& ar, const u8* c1, int m, const u8* c2, int n)
{
	w.c1 = c1; w.c2 = c2; w.m = m; w.n = n; w.mark = ar.used; w.fmark = fast.used;
	w.code2 = (u8*)kb_alloc2(fast, ar, (u64)n + 1);
	w.stride = ((u32)n + 3) >> 2;
	w.tb = (u8*)kb_alloc2(fast, ar, (u64)w.stride * (u64)m);
	for (int k = 0; k < 2; k++)
	{
		w.bS[k] = m > 32 ? (int*)kb_alloc2(fast, ar, (u64)(n + 1) * 4) : nullptr;
		w.bT[k] = m > 32 ? (int*)kb_alloc2(fast, ar, (u64)(n + 1) * 4) : nullptr;
	}
	w.i0 = 0; w.cur = 0; w.h = m < 32 ? m : 32;
	return !ar.ovf;
}

// all lanes: the codes of c2 (row 0 of the DP is analytic: S[0][j] = -2-j, T[0][j] = -inf)
KB_HD void kb_nww_init_rows(KbNwWarp& w, int t)
{
	for (int j = t + 1; j <= w.n; j += 32) w.code2[j] = (u8)kb_nt4(w.c2[j - 1]);
}

// all lanes: start of a strip (column 0 of the lane's row); caches the strip constants in the lane
KB_HD void kb_nww_strip_begin(KbNwWarp& w, KbNwLane& L, int t)
{
	int i = w.i0 + t + 1;
	L.n = w.n; L.h = w.h; L.last = (t == w.h - 1 && w.i0 + w.h < w.m) ? 1 : 0;
	L.a = t < w.h ? kb_nt4(w.c1[i - 1]) : 4;
	L.left_s = -2 - i; L.left_r = KB_NW_NEG; L.diag = i == 1 ? 0 : -2 - (i - 1); L.pack = 0;
	L.code2 = w.code2; L.tbrow = w.tb + (u64)w.stride * (u64)(w.i0 + t);
	L.rS = w.i0 > 0 ? w.bS[w.cur] : nullptr; L.rT = w.i0 > 0 ? w.bT[w.cur] : nullptr;
	L.wS = w.bS[w.cur ^ 1]; L.wT = w.bT[w.cur ^ 1];
	if (t == 0 && L.wS != nullptr) { L.wS[0] = -2 - (w.i0 + w.h); L.wT[0] = -2 - (w.i0 + w.h); }
}

// all lanes: one anti-diagonal step
KB_HD void kb_nww_step(KbNwWarp& w, KbNwLane& L, int d, int t)
{
	int j = d - t + 1;
	if (t >= L.h || j < 1 || j > L.n) return;
	int us, ut;
	if (t == 0) { if (L.rS != nullptr) { us = L.rS[j]; ut = L.rT[j]; } else { us = -2 - j; ut = KB_NW_NEG; } }
	else { us = w.xs[(d - 1) & 1][t - 1]; ut = w.xt[(d - 1) & 1][t - 1]; }
	int r = L.left_r - 1 > L.left_s - 3 ? L.left_r - 1 : L.left_s - 3;
	int tt = ut - 1 > us - 3 ? ut - 1 : us - 3;
	int dg = L.diag + (L.a == (int)L.code2[j] ? 3 : -3);
	int s = dg > r ? dg : r; if (tt > s) s = tt;
	L.diag = us; L.left_s = s; L.left_r = r;
	w.xs[d & 1][t] = s; w.xt[d & 1][t] = tt;
	if (L.last) { L.wS[j] = s; L.wT[j] = tt; }
	u32 bits = (s == r ? 1u : 0u) | (s == tt ? 2u : 0u);
	int q = (j - 1) & 3;
	L.pack |= bits << (q << 1);
	if (q == 3 || j == L.n) { L.tbrow[(j - 1) >> 2] = (u8)L.pack; L.pack = 0; }
}

// lane 0: next strip; returns false when all rows are done
KB_HD bool kb_nww_strip_end(KbNwWarp& w)
{
	w.i0 += 32; w.cur ^= 1;
	if (w.i0 >= w.m) return false;
	w.h = w.m - w.i0 < 32 ? w.m - w.i0 : 32;
	return true;
}

// lane 0: traceback (nw_alignment.cpp:59-72): gap-in-read first, then gap-in-genome, else diagonal. The runs are written
// backwards from out[cap) (cap >= m + n), so they end up in alignment order at out[first .. cap); returns first.
KB_HD int kb_nww_traceback(KbNwWarp& w, KbArena& fast, KbArena& ar, u32* out, int cap, int* ident_out, int* aligned_out)
{
	int m = w.m, n = w.n;
	int i = m, j = n, wr = cap, ident = 0, aligned = 0, cur = -1, len = 0;
	while (i > 0 || j > 0)
	{
		int type;
		if (i == 0) type = KB_RUN_D;
		else if (j == 0) type = KB_RUN_I;
		else
		{
			int bits = (w.tb[(u64)w.stride * (u64)(i - 1) + (u64)((j - 1) >> 2)] >> (((j - 1) & 3) << 1)) & 3;
			type = (bits & 1) ? KB_RUN_D : ((bits & 2) ? KB_RUN_I : KB_RUN_M);
		}
		if (type == KB_RUN_D) j--;
		else if (type == KB_RUN_I) i--;
		else { i--; j--; aligned++; if (w.c1[i] == w.c2[j]) ident++; }
		if (type == cur) len++;
		else { if (len > 0) out[--wr] = ((u32)len << 2) | (u32)cur; cur = type; len = 1; }
	}
	if (len > 0) out[--wr] = ((u32)len << 2) | (u32)cur;
	*ident_out = ident; *aligned_out = aligned;
	ar.used = w.mark; fast.used = w.fmark;
	return wr;
}

// ================================================================================================
// GenerateNormalPairAlignment as a resumable iterator: optional 8-mer partition, pure-gap / copy pieces are appended to
// the run list directly, every piece that needs nw_alignment is handed back to the caller (the warp).
// The reference recurses (tools.cpp:197, pacbio pieces > 300); here the recursion is an explicit depth-first work stack
// so that pieces are appended to `acc` in exactly the reference's left-to-right order.
// ================================================================================================
enum { KB_W_FRAG = 0, KB_W_NW = 1, KB_W_COPY = 2, KB_W_INS = 3, KB_W_DEL = 4 };
struct KbWork { i32 r0, rl, g0, gl, kind, pad; };
struct KbWorkP { i32 r0, rl, g0; u32 glk; };   // the work stack's 16-byte form: gl in the low 28 bits, kind above
KB_HD KbWorkP kb_work_pack(const KbWork& w) { KbWorkP p; p.r0 = w.r0; p.rl = w.rl; p.g0 = w.g0; p.glk = (u32)w.gl | ((u32)w.kind << 28); return p; }
KB_HD KbWork kb_work_unpack(const KbWorkP& p) { KbWork w; w.r0 = p.r0; w.rl = p.rl; w.g0 = p.g0; w.gl = (i32)(p.glk & 0x0FFFFFFFu); w.kind = (i32)(p.glk >> 28); w.pad = 0; return w; }

// nw_alignment size class of an (rl x gl) problem at text position g
KB_HD u32 kb_piece_class(const KbIndexDev& ix, const KbBatchDev& bt, int rl, int gl, i64 g)
{
	const int mx = rl > gl ? rl : gl;
	if (mx > bt.nw_tmax || g < 0 || g + gl > ix.G2) return (u32)(KB_NW_CLASSES - 1);   // too large for one thread, or touching the outside of the text
	return mx <= 32 ? (u32)((mx - 1) >> 3) : (mx <= 64 ? 4u : 5u);
}
// registers one nw_alignment problem; false when the piece arena is full (flagged)
KB_HD bool kb_emit_piece(const KbIndexDev& ix, const KbBatchDev& bt, u32 job, i64 job_gpos, int r0, int rl, int g0, int gl, u32 out_off, u32 whole)
{
	const u32 id = KB_ALLOC(&bt.counters[24], 1u);
	if (id >= bt.cap_pieces) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_JOBS); return false; }
	KbPiece pc; pc.job = job; pc.r0 = r0; pc.rl = rl; pc.g0 = g0; pc.gl = gl; pc.out_off = out_off; pc.whole = whole; pc.pad = 0;
	bt.pieces[id] = pc;
	const u32 cls = kb_piece_class(ix, bt, rl, gl, job_gpos + g0);
	const u32 slot = KB_ALLOC_KEYED(&bt.counters[16 + cls], 1u);
	bt.piece_list[(size_t)cls * bt.cap_pieces + slot] = id;
	return true;
}

// Where the pieces of a partitioned job go: the job's slice of the run arena (zeroed beforehand), filled left to right.
// Every piece owns max(1, rl + gl) consecutive entries: literal pieces (pure gaps, copies) store their single run there,
// nw_alignment pieces are registered and their solver writes into the same entries later; k_align_gather then squeezes
// the zeros out and merges equal neighbours, which is what the reference's sequential AddNewCigarElements calls produce.
struct KbSink
{
	const KbIndexDev* ix; const KbBatchDev* bt; u32 job; i64 gpos; u32 base, cap, cur; int ident, aligned; bool ovf;
	KB_HD void lit(int type, int len, int span)
	{
		if (len <= 0) return;
		if (cur + (u32)span > cap) { ovf = true; return; }
		bt->runs[base + cur] = ((u32)len << 2) | (u32)type; cur += (u32)span;
	}
	KB_HD void piece(int r0, int rl, int g0, int gl)
	{
		if (cur + (u32)(rl + gl) > cap) { ovf = true; return; }
		if (!kb_emit_piece(*ix, *bt, job, gpos, r0, rl, g0, gl, base + cur, 0u)) ovf = true;
		cur += (u32)(rl + gl);
	}
};

KB_HD int kb_mismatch_packed(const KbIndexDev& ix, const KbPk* rd, const u8* f1, int rpos, i64 gpos, int n, int limit);
KB_HD u8 kb_ref_char(const KbIndexDev& ix, i64 p);
// 32 bases of a packed read from position pos on (codes only)
KB_HD u64 kb_read_code(const KbPk* rd, int pos)
{
	const u64 a = rd[pos >> 5].code; const int s = (pos & 31) * 2;
	return s ? (a << s) | (rd[(pos >> 5) + 1].code >> (64 - s)) : a;
}
// Where do eight equal bases in a row start? a, b: 32 bases each (base i at bits 63-2i, 62-2i); bit 62-2i of the result is set iff
// bases i .. i+7 are pairwise equal (so only for i <= 24): exactly "the 8-mer ids at the two positions are equal" for pure-base text.
KB_HD u64 kb_match8(u64 a, u64 b)
{
	const u64 x = a ^ b; u64 e = ~(x | (x >> 1)) & 0x5555555555555555ull;
	e &= e << 2; e &= e << 4; e &= e << 8;
	return e;
}
struct KbFragIter
{
	const KbParams* pm; KbArena* ar; KbArena* fast; const u8* f1; const u8* f2; KbSink* sink;
	// the job's place in the packed read and in the text (set by kb_pt_begin): fragments of pure bases are partitioned on the packed
	// words (part_pairs_packed), nothing of them is copied; f1 points at the read's characters in HBM, f2 (reference characters) is
	// only materialised for a fragment that holds a character which is no base
	const KbIndexDev* ix; const KbPk* rd; int rbase, gl0; i64 gbase; u8* f2w;
	// Work stack: worst case rl0 + gl0 + 4 entries, in practice a handful. The first `sfast` entries are tried in the warp's shared-memory
	// pool, the rest lives in the HBM arena (r23: with the whole stack in front of them, the 8-mer id arrays of every fragment beyond
	// ~100 x 100 ended up in HBM, where part_pairs reads them n1 x (2 shift + 1) times).
	KbWorkP* st; KbWorkP* st2; int sp, scap, sfast;
	KB_HD KbWorkP& slot(int i) { return i < sfast ? st[i] : st2[i - sfast]; }
	// the fragment being partitioned (between next() == 2 and part_finish())
	KbWork cur; u32* w1; u32* w2; KbSeg* raw; int shift, cap, cap_full, dirty; u32 np; u64 pmark, fmark;

	// fast_: the warp's shared-memory pool, used for whatever fits; ar_: its arena in HBM
	KB_HD bool init(const KbParams* pm_, KbArena* fast_, KbArena* ar_, const u8* f1_, int rl0, const u8* f2_, int gl0_, KbSink* sink_)
	{
		pm = pm_; ar = ar_; fast = fast_; f1 = f1_; f2 = f2_; sink = sink_; sp = 0; gl0 = gl0_; f2w = nullptr;
		scap = rl0 + gl0 + 4;
		sfast = sink_->bt->part_stack > 0 && sink_->bt->part_stack < scap ? sink_->bt->part_stack : scap;
		st = (KbWorkP*)kb_alloc2(*fast, *ar, (u64)sfast * sizeof(KbWorkP));
		st2 = scap > sfast ? (KbWorkP*)ar->alloc((u64)(scap - sfast) * sizeof(KbWorkP)) : nullptr;
		if (st == nullptr || (scap > sfast && st2 == nullptr)) return false;
		KbWork w; w.r0 = 0; w.rl = rl0; w.g0 = 0; w.gl = gl0; w.kind = KB_W_FRAG; w.pad = 0; slot(sp++) = kb_work_pack(w);
		return true;
	}

	// one lane. 2: a fragment is ready to be partitioned (part_scan, part_ids, part_pairs by all lanes with a barrier after
	// each, then part_finish by one lane); 0: finished (or something overflowed). Everything else is written to the sink.
	KB_HD int next()
	{
		while (sp > 0 && !ar->ovf && !sink->ovf)
		{
			const KbWork e = kb_work_unpack(slot(--sp));
			if (e.kind == KB_W_INS) { sink->lit(KB_RUN_I, e.rl, e.rl); continue; }
			if (e.kind == KB_W_DEL) { sink->lit(KB_RUN_D, e.gl, e.gl); continue; }
			if (e.kind == KB_W_COPY)
			{
				// identical characters (raw equality against the upper-case reference, tools.cpp:84): 32 per step on the packed words
				const int id = e.rl - kb_mismatch_packed(*ix, rd, f1 + e.r0, rbase + e.r0, gbase + e.g0, e.rl, 0x3FFFFFFF);
				sink->lit(KB_RUN_M, e.rl, e.rl + e.gl); sink->ident += id; sink->aligned += e.rl;
				continue;
			}
			if (e.kind == KB_W_FRAG && e.rl > 30 && e.gl > 30)
			{
				int rl = e.rl, gl = e.gl;
				if (pm->pacbio) { shift = rl > gl ? (int)(rl * 0.2) : (int)(gl * 0.2); if (shift > 50) shift = 50; }
				else shift = pm->max_gaps;
				pmark = ar->used; fmark = fast->used;
				w1 = nullptr; w2 = nullptr;   // only a fragment with a character that is no base needs the id arrays (part_ids)
				cap_full = ((rl < gl ? rl : gl) / 9 + 2) * (2 * shift + 1);   // exact-match runs on one diagonal start >= 9 apart
				if (cap_full > rl + gl) cap_full = rl + gl;
				// that bound is hundreds of entries, a fragment usually has a handful of runs: a short list (in the pool when it fits) first,
				// part_grow() switches to the full one in HBM when the runs did not fit
				cap = sink->bt->part_raw > 0 && sink->bt->part_raw < cap_full ? sink->bt->part_raw : cap_full;
				raw = (KbSeg*)kb_alloc2(*fast, *ar, (u64)cap * sizeof(KbSeg));
				if (ar->ovf) return 0;
				cur = e; np = 0; dirty = 0;
				return 2;
			}
			sink->piece(e.r0, e.rl, e.g0, e.gl);
		}
		return 0;
	}

	// all lanes: does either side hold a character that is no base? (then the literal scan of kb_kmer_ids is replayed on characters)
	KB_HD void part_scan(int lane)
	{
		int bad = 0;
		for (int o = 32 * lane; o < cur.rl; o += 32 * 32)
		{
			const int m = cur.rl - o < 32 ? cur.rl - o : 32;
			if (kb_read_win(rd, rbase + cur.r0 + o).n4 & ~((~0u >> (m - 1)) >> 1)) bad = 1;
		}
		for (int o = 32 * lane; o < cur.gl; o += 32 * 32)
		{
			const int m = cur.gl - o < 32 ? cur.gl - o : 32; u32 inv;
			kb_ref_win(*ix, gbase + cur.g0 + o, &inv);
			if (inv & ~((~0u >> (m - 1)) >> 1)) bad = 1;
		}
		if (bad) dirty = 1;
	}
	// one lane, dirty fragments only: 8-mer ids of both sides by the reference's literal scan (the reference characters are fetched now)
	KB_HD void part_ids(int lane)
	{
		if (!dirty || lane != 0) return;
		w1 = (u32*)kb_alloc2(*fast, *ar, (u64)cur.rl * 4); w2 = (u32*)kb_alloc2(*fast, *ar, (u64)cur.gl * 4);
		u8* b = (u8*)ar->alloc((u64)cur.gl);
		if (ar->ovf) return;
		for (int i = 0; i < cur.gl; i++) b[i] = kb_ref_char(*ix, gbase + cur.g0 + i);
		kb_kmer_ids(cur.rl, f1 + cur.r0, w1); kb_kmer_ids(cur.gl, b, w2);
	}
	// one exact-match run found: rpos / gpos relative to the fragment, l bases
	KB_HD void part_emit(int r, int g, int l)
	{
		const u32 slot = KB_ATOMIC_ADD(&np, 1u);
		if ((int)slot < cap) { KbSeg sg; sg.simple = 1; sg.rpos = r; sg.gpos = (i64)g; sg.rlen = sg.glen = l; raw[slot] = sg; }
	}
	// all lanes, pure-base fragments: the same runs from the packed words. A run of equal 8-mer ids along a diagonal is a maximal run
	// of L >= 8 equal bases (L - 7 consecutive id pairs, reported length 8 + (L - 7) - 1 = L). The id pairs of a diagonal are cut into
	// cells of 25 positions: kb_match8 marks where eight equal bases start; a cell reports the runs that START in it (its first
	// position starts one unless the base in front is equal too) and measures them forward with XOR + count-leading-zeros, across
	// cell borders. Cells are independent, so lanes take them in any order: P (a power of two) lanes per diagonal, 32 / P diagonals
	// per round. r24/r26 (C5): comparing ids cell by cell was a third of k_align_part -- 46 G (position, diagonal) cells per 50 k reads.
	KB_HD void part_pairs_packed(int lane)
	{
		const u64 M5 = 0x5555555555555555ull;
		const int rl = cur.rl, gl = cur.gl, n1 = rl - 7, n2 = gl - 7;
		if (n1 <= 0 || n2 <= 0) return;
		int dlo = -(n1 - 1), dhi = n2 - 1;
		if (dlo < -(shift - 1)) dlo = -(shift - 1);
		if (dhi > shift - 1) dhi = shift - 1;
		const int rp0 = rbase + cur.r0; const i64 gp0 = gbase + cur.g0;
		const int longest = n1 < n2 ? n1 : n2, ncell_max = (longest + 24) / 25;
		int lg = 0; while ((1 << lg) < ncell_max && lg < 5) lg++;
		const int P = 1 << lg, dper = 32 >> lg;
		for (int dbase = dlo; dbase <= dhi; dbase += dper)
		{
			const int d = dbase + (lane >> lg);
			if (d > dhi) continue;
			const int r_lo = d < 0 ? -d : 0, r_end = n1 < n2 - d ? n1 : n2 - d;   // id pairs (r, r + d) exist for r in [r_lo, r_end)
			for (int rs = r_lo + 25 * (lane & (P - 1)); rs < r_end; rs += 25 * P)
			{
				u32 inv;
				u64 S = kb_match8(kb_read_code(rd, rp0 + rs), kb_ref_win(*ix, gp0 + rs + d, &inv));
				const int cnt = r_end - rs < 25 ? r_end - rs : 25;
				S &= ~(~0ull >> (2 * cnt));                                         // the cell's own positions
				if (S == 0) continue;
				u64 st = S & ~(S >> 2);                                             // a position whose predecessor is no match
				if ((st >> 62) & 1ull)                                              // the cell's first position: is the base in front equal as well?
				{
					if (rs > r_lo && (((kb_read_code(rd, rp0 + rs - 1) ^ kb_ref_win(*ix, gp0 + rs - 1 + d, &inv)) >> 62) & 3ull) == 0ull) st &= ~(1ull << 62);
				}
				while (st)
				{
					const int i = (int)KB_CLZLL(st) >> 1; st &= ~(1ull << (62 - 2 * i));
					const int r = rs + i, g = r + d, lim = rl - r < gl - g ? rl - r : gl - g;
					int l = 0;
					while (l < lim)
					{
						u64 x = kb_read_code(rd, rp0 + r + l) ^ kb_ref_win(*ix, gp0 + g + l, &inv); x = (x | (x >> 1)) & M5;
						const int same = x ? (int)KB_CLZLL(x) >> 1 : 32;
						l += same;
						if (same < 32) break;
					}
					if (l > lim) l = lim;
					part_emit(r, g, l);
				}
			}
		}
	}
	// all lanes: the exact-match runs of kb_kmer_pairs (min_len 8), appended in arbitrary order (part_finish sorts them by a total
	// order). A lane takes read positions r = lane, lane + 32, ... and walks the diagonals |g - r| < shift of each: the id of r stays in
	// a register, the ids of g are consecutive loads (r24, C5: dealing (position, diagonal) cells to the lanes through idx / nd and
	// idx % nd spent 37 % of the kernel's instructions on the division and reloaded w1[r] for every cell).
	KB_HD void part_pairs(int lane)
	{
		if (!dirty) { part_pairs_packed(lane); return; }
		const int n1 = cur.rl - 7, n2 = cur.gl - 7;
		if (n1 <= 0 || n2 <= 0 || w1 == nullptr || w2 == nullptr) return;
		for (int r = lane; r < n1; r += 32)
		{
			const u32 id = w1[r];
			if (id == KB_NOKMER) continue;
			const int glo = r - (shift - 1) > 0 ? r - (shift - 1) : 0, ghi = r + (shift - 1) < n2 - 1 ? r + (shift - 1) : n2 - 1;
			const u32 before = r > 0 ? w1[r - 1] : KB_NOKMER;
			for (int g = glo; g <= ghi; g++)
			{
				if (w2[g] != id) continue;
				if (g > 0 && before != KB_NOKMER && before == w2[g - 1]) continue;   // not the start of its run
				int run = 1;
				while (r + run < n1 && g + run < n2 && w1[r + run] != KB_NOKMER && w1[r + run] == w2[g + run]) run++;
				part_emit(r, g, 8 + run - 1);
			}
		}
	}
	// one lane, after part_pairs: true when the runs did not fit the short list -- the full-size list is then in place (HBM arena) and
	// part_pairs has to run again
	KB_HD bool part_grow()
	{
		if ((int)np <= cap || cap >= cap_full) return false;
		raw = (KbSeg*)ar->alloc((u64)cap_full * sizeof(KbSeg));
		if (ar->ovf) return false;
		cap = cap_full; np = 0;
		return true;
	}
	// one lane: IdentifyNormalPairs on the runs; pushes the pieces, or registers the whole fragment as one nw_alignment problem
	KB_HD void part_finish()
	{
		const KbWork e = cur; const int rl = e.rl, gl = e.gl;
		if ((int)np > cap) { ar->ovf = true; return; }
		int tot = 0; KbSeg* part = nullptr; const int n = (int)np;
		if (n > 0)
		{
			kb_sort_segs<true>(raw, n);
			part = (KbSeg*)kb_alloc2(*fast, *ar, (u64)(2 * n + 2) * sizeof(KbSeg));
			i32* order = (i32*)kb_alloc2(*fast, *ar, (u64)n * 4);
			if (ar->ovf) return;
			tot = kb_fill_pairs(rl, gl, raw, n, part, order);
		}
		if (tot > 0)
		{
			if (sp + tot > scap) { ar->ovf = true; return; }
			for (int i = tot - 1; i >= 0; i--)   // reversed, so that pops come out left to right
			{
				const KbSeg p = part[i];
				if (p.rlen <= 0 && p.glen <= 0) continue;
				KbWork w; w.r0 = e.r0 + p.rpos; w.rl = p.rlen; w.g0 = e.g0 + (i32)p.gpos; w.gl = p.glen; w.pad = 0;
				if (p.glen == 0) w.kind = KB_W_INS;
				else if (p.rlen == 0) w.kind = KB_W_DEL;
				else if ((p.rlen == 1 && p.glen == 1) || p.simple) w.kind = KB_W_COPY;
				else if (pm->pacbio && (p.rlen > 300 || p.glen > 300)) w.kind = KB_W_FRAG;
				else w.kind = KB_W_NW;
				slot(sp++) = kb_work_pack(w);
			}
			ar->used = pmark; fast->used = fmark;
			return;
		}
		ar->used = pmark; fast->used = fmark;
		sink->piece(e.r0, e.rl, e.g0, e.gl);
	}
};

// ================================================================================================
// Fragment-pair processing, split in three phases so that the expensive, divergent part (partition + NW) runs in its own
// kernel over a compact job list:
//   phase A  kb_segments_read : IdentifyNormalPairs per surviving candidate, every segment classified; the quick tests of
//                                Process{Normal,Head,Tail}SequencePair (tools.cpp:229-244,301-311,352-362) are decided here
//   phase B  kb_align_job     : GenerateNormalPairAlignment (tools.cpp:142) for the segments that need it
//   phase C  kb_assemble_read : cigar elements, head/tail post-processing, GapPenalty, coordinates, best/sub-score
// ================================================================================================
KB_HD bool kb_same_chromosome(const KbIndexDev& ix, const KbSeg* v, int n)   // CheckCoordinateValidity :582
{
	i64 a = 0, b = ix.G2;
	for (int i = 0; i < n; i++) if (v[i].glen > 0) { a = v[i].gpos; break; }
	for (int i = n - 1; i >= 0; i--) if (v[i].glen > 0) { b = v[i].gpos + v[i].glen - 1; break; }
	if ((a < ix.G) != (b < ix.G)) return false;
	int ea = kb_chr_lookup(ix, a), eb = kb_chr_lookup(ix, b);
	return ea < ix.n_ends && eb < ix.n_ends && ix.end_chr[ea] == ix.end_chr[eb];
}

KB_HD u8 kb_ref_char(const KbIndexDev& ix, i64 p) { return kb_code_char(kb_ref_code(ix, p)); }

// mismatches between read[rpos..] and the reference at gpos, stopping once more than `limit` were seen
KB_HD int kb_mismatch_ref(const KbIndexDev& ix, const u8* f1, i64 gpos, int n, int limit)
{
	int c = 0;
	for (int i = 0; i < n && c <= limit; i++) if (f1[i] != kb_ref_char(ix, gpos + i)) c++;
	return c;
}

// the same count, 32 characters per step on the packed read and the 64-bit reference words; a chunk that holds a character
// other than upper-case ACGT on either side is compared character by character (raw char equality, tools.cpp:44)
KB_HD int kb_mismatch_packed(const KbIndexDev& ix, const KbPk* rd, const u8* f1, int rpos, i64 gpos, int n, int limit)
{
	const u64 M5 = 0x5555555555555555ull;
	int c = 0;
	for (int o = 0; o < n && c <= limit; o += 32)
	{
		const int m = n - o < 32 ? n - o : 32;
		KbPk rw = kb_read_win(rd, rpos + o); u32 ginv; u64 gw = kb_ref_win(ix, gpos + o, &ginv);
		const u32 lm = ~((~0u >> (m - 1)) >> 1);   // first m of 32
		if (((rw.bad | ginv) & lm) == 0)
		{
			u64 x = rw.code ^ gw; x = (x | (x >> 1)) & M5;
			x &= ~(((~0ull >> (m - 1)) >> (m - 1)) >> 2);   // first m of 32 two-bit fields
			c += (int)KB_POPCLL(x);
		}
		else for (int i = 0; i < m; i++) if (f1[o + i] != kb_ref_char(ix, gpos + o + i)) c++;
	}
	return c;
}

// quick test of tools.cpp:240/301/352: equal length, <= 2 mismatches and <= 20 %
KB_HD bool kb_quick_match_ref(const KbIndexDev& ix, const KbPk* rd, const u8* f1, const KbSeg& sp, int* n)
{
	if (sp.rlen != sp.glen) return false;
	*n = kb_mismatch_packed(ix, rd, f1, sp.rpos, sp.gpos, sp.rlen, 2);
	return *n <= 2 && *n <= (int)(sp.rlen * 0.2);
}

// registers the fragment pair `sp` of read r as an alignment job (GenerateNormalPairAlignment, tools.cpp:142) and routes it:
// fragments with both sides > 30 go through the 8-mer partition first (tools.cpp:146), everything else is one nw_alignment
// problem, registered right away. whole: one nw_alignment problem whatever the size (the stage-level test entry kb_debug_align).
KB_HD void kb_make_job(const KbIndexDev& ix, const KbBatchDev& bt, int r, const KbSeg& sp, bool whole, KbSegX* out)
{
	// one 64-bit atomic reserves the job id (high word = counters[11]) and its slice of the run arena (low word = counters[10])
	u32 need = (u32)(sp.rlen + sp.glen + 2);
	unsigned long long old = KB_ALLOC(reinterpret_cast<unsigned long long*>(bt.counters + 10), (unsigned long long)((1ull << 32) | (unsigned long long)need));
	u32 id = (u32)(old >> 32), ro = (u32)old;
	if (id >= bt.cap_jobs) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_JOBS); return; }
	if ((u64)ro + need > (u64)bt.cap_runs) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_RUNS); return; }
	KbJob jb; jb.gpos = sp.gpos; jb.read = (u32)r; jb.rpos = sp.rpos; jb.rlen = sp.rlen; jb.glen = sp.glen; jb.run_off = ro; jb.nruns = 0; jb.ident = 0; jb.aligned = 0;
	bt.jobs[id] = jb;
	if (!whole && sp.rlen > 30 && sp.glen > 30) { const u32 slot = KB_ALLOC(&bt.counters[23], 1u); bt.part_list[slot] = id; }
	else kb_emit_piece(ix, bt, id, sp.gpos, 0, sp.rlen, 0, sp.glen, ro, 1u);
	out->info = KB_SEG_JOB; out->aux = id;
}

KB_HD void kb_classify_segment(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, const u8* seq, const KbPk* rd, const KbSeg& sp, int j, int n, KbSegX* out)
{
	out->s = sp; out->info = KB_SEG_SKIP; out->aux = 0;
	if (sp.rlen == 0 && sp.glen == 0) return;
	if (sp.simple) { out->info = KB_SEG_SIMPLE; return; }
	const u8* f1 = seq + sp.rpos; int mm;
	if (j == 0 || j == n - 1)
	{
		if (sp.rlen > 3000) { out->info = KB_SEG_SOFT; return; }                                  // AlignmentCandidates.cpp:671,690
		if (!pm.pacbio && kb_quick_match_ref(ix, rd, f1, sp, &mm)) { out->info = KB_SEG_QUICK; out->aux = (u32)(sp.rlen - mm); return; }
		if (!pm.pacbio && sp.rlen > (j == 0 ? 50 : 100)) { out->info = KB_SEG_SOFT; return; }     // tools.cpp:307,358
	}
	else
	{
		if (sp.rlen == 0 || sp.glen == 0) { out->info = KB_SEG_GAP; return; }
		if (kb_quick_match_ref(ix, rd, f1, sp, &mm)) { out->info = KB_SEG_QUICK; out->aux = (u32)(sp.rlen - mm); return; }
	}
	if (sp.rlen == 1 && sp.glen == 1)
	{
		// nw_alignment on 1 x 1 always yields one aligned column (diag +-3 beats both gap paths at -6)
		out->info = KB_SEG_ONE; out->aux = f1[0] == kb_ref_char(ix, sp.gpos) ? 1u : 0u;
		return;
	}
	kb_make_job(ix, bt, r, sp, false, out);
}

// phase A
#define KB_SEG_FAST 8   // candidates with at most this many seeds are expanded in local memory instead of the HBM arena
// Classified-segment slots. Lanes arrive here one or two at a time, so warp aggregation does not help and the grid-wide cursor
// takes one atomic per candidate (ncu r16: the allocation sites are ~45 % of k_segments' stall samples). With KB_SEG_SLAB the
// warp reserves a range up front (one global atomic) and its lanes take slots from it through shared memory; a warp that runs
// out falls back to the cursor. Holes are harmless: every consumer goes through cseg_off / cseg_n. 256 slots per warp by default
// (r35 A/B at C3: k_segments 1.64 -> 1.41 ms per 2.5 M reads, identical results).
KB_HD u32 kb_alloc_segx(const KbBatchDev& bt, u32 n)
{
#if defined(__CUDA_ARCH__)
	if (bt.segx_slab != nullptr)
	{
		const u32 at = atomicAdd(&bt.segx_slab[0], n);
		if (at + n <= bt.segx_slab[1]) return at;
	}
#endif
	return KB_ALLOC(&bt.counters[8], n);
}
KB_HD void kb_segments_cand(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, const u8* seq, const KbPk* rd, int rlen,
                            u32 ci, const KbCand& c, KbSeg* in, KbSeg* sv, i32* order)
{
	const int ns = c.nseg;
	for (int k = 0; k < ns; k++) in[k] = bt.segs[c.seg_start + k];
	int n = kb_fill_pairs(rlen, -1, in, ns, sv, order);
	if (!kb_same_chromosome(ix, sv, n)) return;
	u32 off = kb_alloc_segx(bt, (u32)n);
	if ((u64)off + (u64)n > (u64)bt.cap_segx) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEGX); return; }
	bt.cseg_off[ci] = off; bt.cseg_n[ci] = n;
	for (int j = 0; j < n; j++) kb_classify_segment(ix, pm, bt, r, seq, rd, sv[j], j, n, &bt.segx[off + j]);
}

// The common case needs no staging at all: when the (gPos,rPos)-sorted seeds of a candidate neither overlap nor invert (each
// ends before the next begins, in the read and in the genome), RemoveTandemRepeatSeeds, RemoveTranslocatedSeeds and
// CheckOverlappingSeeds change nothing and the stable merge of :449 simply interleaves seed, gap, seed, ... So the segment
// list is streamed straight from the seed array: one pass to count and validate, one pass to classify and store.
// Returns false (nothing written) when the seeds need the general path.
KB_HD bool kb_segments_cand_stream(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, const u8* seq, const KbPk* rd, int rlen, u32 ci, const KbCand& c)
{
	const KbSeg* in = bt.segs + c.seg_start; const int ns = c.nseg;
	KbSeg prev = in[0]; const KbSeg first = prev;
	int n = ns;
	for (int k = 1; k < ns; k++)
	{
		const KbSeg cur = in[k];
		const int rg = cur.rpos - (prev.rpos + prev.rlen); const i64 gg = cur.gpos - (prev.gpos + prev.glen);
		if (rg < 0 || gg < 0) return false;
		if (rg > 0 || gg > 0) n++;
		prev = cur;
	}
	const int head = first.rpos > 0 ? first.rpos : 0, tail = rlen - (prev.rpos + prev.rlen);
	if (head > 0) n++;
	if (tail > 0) n++;
	// CheckCoordinateValidity :582 on the first and last segment with a genome side
	{
		i64 a = first.gpos, b = prev.gpos + prev.glen - 1;
		if (head > 0) { a = first.gpos - head; if (a < 0) a = 0; }
		if (tail > 0) b = prev.gpos + prev.glen + tail - 1;
		if ((a < ix.G) != (b < ix.G)) return true;
		int ea = kb_chr_lookup(ix, a), eb = kb_chr_lookup(ix, b);
		if (!(ea < ix.n_ends && eb < ix.n_ends && ix.end_chr[ea] == ix.end_chr[eb])) return true;
	}
	u32 off = kb_alloc_segx(bt, (u32)n);
	if ((u64)off + (u64)n > (u64)bt.cap_segx) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEGX); return true; }
	bt.cseg_off[ci] = off; bt.cseg_n[ci] = n;
	// one loop, one call site of kb_classify_segment: the lanes of a warp then classify their j-th segments together,
	// whatever kind they are (four call sites would serialise heads, seeds, gaps and tails behind one another)
	int k = 0; bool gap_done = false;
	prev = first;
	for (int j = 0; j < n; j++)
	{
		KbSeg sg;
		if (j == 0 && head > 0) { sg.simple = 0; sg.rpos = 0; sg.gpos = first.gpos - head; if (sg.gpos < 0) sg.gpos = 0; sg.rlen = head; sg.glen = head; }
		else if (k >= ns) { sg.simple = 0; sg.rpos = prev.rpos + prev.rlen; sg.gpos = prev.gpos + prev.glen; sg.rlen = tail; sg.glen = tail; }
		else
		{
			const KbSeg cur = in[k];
			const int rg = k > 0 ? cur.rpos - (prev.rpos + prev.rlen) : 0, gg = k > 0 ? (int)(cur.gpos - (prev.gpos + prev.glen)) : 0;
			if (!gap_done && (rg > 0 || gg > 0)) { sg.simple = 0; sg.rpos = prev.rpos + prev.rlen; sg.gpos = prev.gpos + prev.glen; sg.rlen = rg; sg.glen = gg; gap_done = true; }
			else { sg = cur; prev = cur; k++; gap_done = false; }
		}
		kb_classify_segment(ix, pm, bt, r, seq, rd, sg, j, n, &bt.segx[off + j]);
	}
	return true;
}

// ar == nullptr: local memory only; returns false (nothing written) when a candidate needs the arena
KB_HD bool kb_segments_read(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, KbArena* ar)
{
	KbCand* cv = bt.cands + bt.cand_off[r];
	int ncan = bt.n_cands[r];
	const u8* seq = bt.seq + bt.seq_off[r]; int rlen = (int)(bt.seq_off[r + 1] - bt.seq_off[r]);
	const KbPk* rd = kb_pk_read(bt, r);
	if (ar == nullptr)
	{
		// all candidates streamable or small? (decided before anything is written: a read is done entirely by one of the two kernels)
		for (int i = 0; i < ncan; i++)
		{
			if (cv[i].score == 0 || cv[i].nseg <= KB_SEG_FAST) continue;
			const KbSeg* in = bt.segs + cv[i].seg_start; bool ok = true;
			for (int k = 1; k < cv[i].nseg && ok; k++) ok = in[k].rpos >= in[k - 1].rpos + in[k - 1].rlen && in[k].gpos >= in[k - 1].gpos + in[k - 1].glen;
			if (!ok) return false;
		}
	}
	KbSeg in_l[KB_SEG_FAST], sv_l[2 * KB_SEG_FAST + 2]; i32 order_l[KB_SEG_FAST];
	for (int i = 0; i < ncan; i++)
	{
		u32 ci = bt.cand_off[r] + (u32)i;
		bt.cseg_n[ci] = -1; bt.cseg_off[ci] = 0;
		const KbCand c = cv[i];
		if (c.score == 0) continue;
		if (kb_segments_cand_stream(ix, pm, bt, r, seq, rd, rlen, ci, c)) continue;
		if (c.nseg <= KB_SEG_FAST) { kb_segments_cand(ix, pm, bt, r, seq, rd, rlen, ci, c, in_l, sv_l, order_l); continue; }
		u64 mark = ar->used;
		int ns = c.nseg;
		KbSeg* in = (KbSeg*)ar->alloc((u64)ns * sizeof(KbSeg));
		KbSeg* sv = (KbSeg*)ar->alloc((u64)(2 * ns + 2) * sizeof(KbSeg));
		i32* order = (i32*)ar->alloc((u64)ns * 4);
		if (ar->ovf) return true;
		kb_segments_cand(ix, pm, bt, r, seq, rd, rlen, ci, c, in, sv, order);
		ar->used = mark;
	}
	return true;
}

// ---- phase B -----------------------------------------------------------------------------------------------------
// B1 k_align_part   : warp per job of the partition list: 8-mer partition (recursive in pacbio mode), literal pieces written,
//                     nw_alignment pieces registered (KbSink / KbFragIter above)
// B2 k_nw_tile<...> : one thread per nw_alignment problem with both sides <= KB_NW_TMAX, by size class (below)
//    k_nw_warp      : one warp per larger problem (anti-diagonal wavefront, kb_nww_*)
// B3 k_align_gather : one thread per partitioned job: squeeze and merge the runs of its pieces
//
// Thread-per-problem solver. The DP is walked in column tiles of TW (<= 32) columns whose S and T values of the previous row
// live in registers (the column loop is fully unrolled), so a cell costs ~15 integer instructions and no memory access;
// per row and tile there is one 64-bit traceback word and, with several tiles, one boundary (S,R) pair carried through
// local memory. Read and reference come as 2-bit packed words (the problem lies inside the text; read characters that are
// no base are flagged in n4/bad and never match, exactly like nst_nt4_table code 4 against a reference base).
// Same recurrence, tie order and run list as kb_nww_*.
#if defined(__CUDA_ARCH__)
#define KB_ADDMAX(a, b, c) __viaddmax_s32((a), (b), (c))   // max(a + b, c): one DPX instruction on sm_100a
#define KB_MAX3(a, b, c) __vimax3_s32((a), (b), (c))
#define KB_UNROLL _Pragma("unroll")
#else
#define KB_ADDMAX(a, b, c) (((a) + (b)) > (c) ? ((a) + (b)) : (c))
#define KB_MAX3(a, b, c) ((a) > (b) ? ((a) > (c) ? (a) : (c)) : ((b) > (c) ? (b) : (c)))
#define KB_UNROLL
#endif

template <int TW, int MAXM, int MAXT>
KB_HD void kb_nwt_solve(const KbIndexDev& ix, const KbBatchDev& bt, const KbPiece& pc, unsigned long long* cells)
{
	const u64 M5 = 0x5555555555555555ull;
	KbJob& jb = bt.jobs[pc.job];
	const int m = pc.rl, n = pc.gl;
	const KbPk* rd = kb_pk_read(bt, (int)jb.read); const int rp = jb.rpos + pc.r0; const i64 gp = jb.gpos + pc.g0;
	const int MAXW = (MAXM + 31) / 32;
	u64 rc[MAXW]; u32 rn4[MAXW], rbad[MAXW]; u64 gc[MAXT];
	for (int w = 0; w < MAXW; w++) if (32 * w < m) { const KbPk k = kb_read_win(rd, rp + 32 * w); rc[w] = k.code; rn4[w] = k.n4; rbad[w] = k.bad; } else { rc[w] = 0; rn4[w] = ~0u; rbad[w] = ~0u; }
	const int nt = (n + 31) >> 5;
	for (int t = 0; t < MAXT; t++) { u32 inv; gc[t] = t < nt ? kb_ref_win(ix, gp + 32 * t, &inv) : 0; }
	u64 tb[MAXM * MAXT];
	int lbs[MAXT > 1 ? MAXM + 1 : 1], lbr[MAXT > 1 ? MAXM + 1 : 1];
	for (int t = 0; t < nt; t++)
	{
		const u64 gw = gc[t]; const int nc = n - 32 * t < TW ? n - 32 * t : TW;
		int S[TW], T[TW];
		KB_UNROLL
		for (int j = 0; j < TW; j++) { S[j] = -2 - (32 * t + j + 1); T[j] = KB_NW_NEG; }
		int prev_old = t == 0 ? 0 : -2 - 32 * t;   // S[i-1][first column of the tile - 1]
		for (int i = 1; i <= m; i++)
		{
			const int w = (i - 1) >> 5, o = (i - 1) & 31;
			const u64 a2 = (rc[w] >> (62 - 2 * o)) & 3ull;
			const u64 x = gw ^ (M5 * a2);
			u64 eq = ~(x | (x >> 1)) & M5;                        // bit 62-2j set: column j of the tile matches the read base
			if ((rn4[w] >> (31 - o)) & 1u) eq = 0;
			int left_s, left_r;
			if (MAXT == 1 || t == 0) { left_s = -2 - i; left_r = KB_NW_NEG; } else { left_s = lbs[i]; left_r = lbr[i]; }
			int diag = prev_old; prev_old = left_s;
			u64 bits = 0;
			KB_UNROLL
			for (int j = 0; j < TW; j++)
			{
				if (j < nc)
				{
					const int us = S[j], ut = T[j];
					const int r = KB_ADDMAX(left_r, -1, left_s - 3);
					const int tt = KB_ADDMAX(ut, -1, us - 3);
					const int dg = diag + (((eq >> (62 - 2 * j)) & 1ull) ? 3 : -3);
					const int sc = KB_MAX3(dg, r, tt);
					bits |= (u64)((sc == r ? 1u : 0u) | (sc == tt ? 2u : 0u)) << (2 * j);
					diag = us; S[j] = sc; T[j] = tt; left_s = sc; left_r = r;
				}
			}
			tb[(i - 1) * MAXT + t] = bits;
			if (MAXT > 1) { lbs[i] = left_s; lbr[i] = left_r; }
		}
	}
	// traceback; runs are written backwards from the end of the piece's slice
	u32* out = bt.runs + pc.out_off; int wr = m + n;
	int i = m, j = n, ident = 0, aligned = 0, cur = -1, len = 0;
	while (i > 0 || j > 0)
	{
		int type;
		if (i == 0) type = KB_RUN_D;
		else if (j == 0) type = KB_RUN_I;
		else { const u32 b2 = (u32)(tb[(i - 1) * MAXT + ((j - 1) >> 5)] >> (2 * ((j - 1) & 31))) & 3u; type = (b2 & 1u) ? KB_RUN_D : ((b2 & 2u) ? KB_RUN_I : KB_RUN_M); }
		if (type == KB_RUN_D) j--;
		else if (type == KB_RUN_I) i--;
		else
		{
			i--; j--; aligned++;
			// raw character equality against the upper-case reference: the read character must be an upper-case base with the same code
			if ((((rbad[i >> 5] >> (31 - (i & 31))) & 1u) == 0u) && (((rc[i >> 5] >> (62 - 2 * (i & 31))) & 3ull) == ((gc[j >> 5] >> (62 - 2 * (j & 31))) & 3ull))) ident++;
		}
		if (type == cur) len++;
		else { if (len > 0) out[--wr] = ((u32)len << 2) | (u32)cur; cur = type; len = 1; }
	}
	if (len > 0) out[--wr] = ((u32)len << 2) | (u32)cur;
	if (pc.whole) { jb.run_off = pc.out_off + (u32)wr; jb.nruns = m + n - wr; jb.ident = ident; jb.aligned = aligned; }
	else { KB_ATOMIC_ADD(&jb.ident, ident); KB_ATOMIC_ADD(&jb.aligned, aligned); }
	*cells += (unsigned long long)m * (unsigned long long)n;
}

// grid-stride loop of one size class
template <int TW, int MAXM, int MAXT>
KB_HD void kb_nwt_class(const KbIndexDev& ix, const KbBatchDev& bt, int cls, u32 first, u32 stride, unsigned long long* cells, unsigned long long* calls)
{
	const u32 count = bt.counters[16 + cls]; const u32* list = bt.piece_list + (size_t)cls * bt.cap_pieces;
	for (u32 q = first; q < count; q += stride) { kb_nwt_solve<TW, MAXM, MAXT>(ix, bt, bt.pieces[list[q]], cells); *calls += 1; }
}

// B2, warp per large problem. State shared by the lanes of the warp (shared memory on the GPU):
struct KbPieceWarp { KbNwWarp nw; KbArena ar, fast; KbPiece pc; u8* f1; u8* f2; const u8* f1g; i64 g; int ok, more_strips; unsigned long long cells; u32 calls; };
// lane 0: open a piece
KB_HD void kb_pw_begin(const KbBatchDev& bt, KbPieceWarp& w, u32 id)
{
	w.pc = bt.pieces[id]; const KbJob& jb = bt.jobs[w.pc.job];
	w.ar.used = 0; w.ar.ovf = false; w.fast.used = 0;
	w.f1 = (u8*)kb_alloc2(w.fast, w.ar, (u64)w.pc.rl); w.f2 = (u8*)kb_alloc2(w.fast, w.ar, (u64)w.pc.gl);
	w.f1g = bt.seq + bt.seq_off[jb.read] + jb.rpos + w.pc.r0; w.g = jb.gpos + w.pc.g0;
	w.ok = (w.f1 != nullptr && w.f2 != nullptr) ? 1 : 0;
	if (w.ok) w.ok = kb_nww_setup(w.nw, w.fast, w.ar, w.f1, w.pc.rl, w.f2, w.pc.gl) ? 1 : 0;
	w.more_strips = 1;
}
// all lanes: characters of both sides into the warp's pool
KB_HD void kb_pw_fetch(const KbIndexDev& ix, KbPieceWarp& w, int t)
{
	if (!w.ok) return;
	for (int i = t; i < w.pc.gl; i += 32) w.f2[i] = kb_ref_char(ix, w.g + i);
	for (int i = t; i < w.pc.rl; i += 32) w.f1[i] = w.f1g[i];
}
// lane 0: traceback and results
KB_HD void kb_pw_end(const KbBatchDev& bt, KbPieceWarp& w)
{
	KbJob& jb = bt.jobs[w.pc.job];
	if (w.ok)
	{
		int ident, aligned; const int cap = w.pc.rl + w.pc.gl;
		const int first = kb_nww_traceback(w.nw, w.fast, w.ar, bt.runs + w.pc.out_off, cap, &ident, &aligned);
		if (w.pc.whole) { jb.run_off = w.pc.out_off + (u32)first; jb.nruns = cap - first; jb.ident = ident; jb.aligned = aligned; }
		else { KB_ATOMIC_ADD(&jb.ident, ident); KB_ATOMIC_ADD(&jb.aligned, aligned); }
		w.calls++; w.cells += (unsigned long long)w.pc.rl * (unsigned long long)w.pc.gl;
	}
	if (w.ar.ovf || !w.ok) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH);
}

// B1, warp per partitioned job
struct KbPartWarp { KbFragIter it; KbSink sink; KbArena ar, fast; u8* f1; u8* f2; const u8* f1g; u32 job; int ok, state, rlen, glen; };
// lane 0: open job `id`
KB_HD void kb_pt_begin(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, KbPartWarp& w, u32 id)
{
	const KbJob& jb = bt.jobs[id];
	w.job = id; w.ar.used = 0; w.ar.ovf = false; w.fast.used = 0; w.rlen = jb.rlen; w.glen = jb.glen;
	w.f1g = bt.seq + bt.seq_off[jb.read] + jb.rpos;
	w.f1 = const_cast<u8*>(w.f1g); w.f2 = nullptr;   // the partition works on the packed read and text (KbFragIter); characters stay where they are
	w.sink.ix = &ix; w.sink.bt = &bt; w.sink.job = id; w.sink.gpos = jb.gpos; w.sink.base = jb.run_off; w.sink.cap = (u32)(jb.rlen + jb.glen + 2); w.sink.cur = 0;
	w.sink.ident = 0; w.sink.aligned = 0; w.sink.ovf = false;
	w.ok = w.it.init(&pm, &w.fast, &w.ar, w.f1, jb.rlen, w.f2, jb.glen, &w.sink) ? 1 : 0;
	w.it.ix = &ix; w.it.rd = kb_pk_read(bt, (int)jb.read); w.it.rbase = jb.rpos; w.it.gbase = jb.gpos;
}
// all lanes: the job's slice of the run arena zeroed
KB_HD void kb_pt_fetch(const KbIndexDev& ix, const KbBatchDev& bt, KbPartWarp& w, int t)
{
	if (!w.ok) return;
	for (u32 i = (u32)t; i < w.sink.cap; i += 32) bt.runs[w.sink.base + i] = 0;
}
// lane 0: close the job (identities of the literal pieces; the nw_alignment pieces add theirs later)
KB_HD void kb_pt_end(const KbBatchDev& bt, KbPartWarp& w)
{
	KbJob& jb = bt.jobs[w.job];
	jb.nruns = 0;
	if (w.sink.ident) KB_ATOMIC_ADD(&jb.ident, w.sink.ident);
	if (w.sink.aligned) KB_ATOMIC_ADD(&jb.aligned, w.sink.aligned);
	if (w.ar.ovf || w.sink.ovf || !w.ok) KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH);
}
// B3: the runs of the job's pieces, in place: zeros squeezed out, equal neighbours merged (= KbRuns::push over the pieces in order)
KB_HD void kb_gather_job(const KbBatchDev& bt, u32 id)
{
	KbJob& jb = bt.jobs[id];
	u32* r = bt.runs + jb.run_off; const int cap = jb.rlen + jb.glen + 2;
	int n = 0; u32 tail = 0;
	for (int k = 0; k < cap; k++)
	{
		const u32 e = r[k]; if (e == 0) continue;
		if (tail != 0 && (tail & 3u) == (e & 3u)) { tail += (e >> 2) << 2; continue; }
		if (tail != 0) r[n++] = tail;
		tail = e;
	}
	if (tail != 0) r[n++] = tail;
	jb.nruns = n;
}

// AddNewCigarElements: run list -> cigar elements (D/I/M)
KB_HD void kb_push_runs(const u32* runs, int first, int last, KbCigar& cg)
{
	for (int k = first; k < last; k++)
	{
		int t = (int)(runs[k] & 3), len = (int)(runs[k] >> 2);
		cg.push(len, t == KB_RUN_D ? KB_OP_D : (t == KB_RUN_I ? KB_OP_I : KB_OP_M));
	}
}

// the alignment of one non-trivial segment as phase C sees it
struct KbAlnView { const u32* runs; int n, ident, aligned; u32 one; };
KB_HD KbAlnView kb_view(const KbBatchDev& bt, const KbSegX& x)
{
	KbAlnView v;
	if (x.info == KB_SEG_ONE) { v.one = (1u << 2) | (u32)KB_RUN_M; v.runs = nullptr; v.n = 1; v.ident = (int)x.aux; v.aligned = 1; }
	else { const KbJob& jb = bt.jobs[x.aux]; v.one = 0; v.runs = bt.runs + jb.run_off; v.n = jb.nruns; v.ident = jb.ident; v.aligned = jb.aligned; }
	return v;
}
KB_HD u32 kb_view_run(const KbAlnView& v, int k) { return v.runs ? v.runs[k] : v.one; }
KB_HD void kb_view_push(const KbAlnView& v, int first, int last, KbCigar& cg)
{
	for (int k = first; k < last; k++)
	{
		u32 e = kb_view_run(v, k); int t = (int)(e & 3), len = (int)(e >> 2);
		cg.push(len, t == KB_RUN_D ? KB_OP_D : (t == KB_RUN_I ? KB_OP_I : KB_OP_M));
	}
}
// CheckLocalAlignmentQuality :255 on the run list
KB_HD bool kb_quality_ok(const KbAlnView& v)
{
	int mis = v.aligned - v.ident;
	return !(v.n >= 4 || (mis >= 3 && mis >= (int)(v.aligned * 0.3)));
}

// ================================================================================================
// Reports
// ================================================================================================
// GenCoordinateInfo + GenerateCIGAR: writes the merged cigar into the global arena
KB_HD void kb_locate_report(const KbIndexDev& ix, const KbBatchDev& bt, bool first, i64 gpos, i64 gend, KbCigar& cg, KbReport& rp)
{
	bool rev = gpos >= ix.G;
	if (!rev)
	{
		rp.fwd = first ? 1 : 0;
		if (ix.n_chr == 1) { rp.chr = 0; rp.pos = gpos + 1; }
		else { int e = kb_chr_lookup(ix, gpos); rp.chr = ix.end_chr[e]; rp.pos = gpos + 1 - ix.chr_fwd[rp.chr]; }
	}
	else
	{
		rp.fwd = first ? 0 : 1;
		if (ix.n_chr == 1) { rp.chr = 0; rp.pos = ix.G2 - gend; }
		else { int e = kb_chr_lookup(ix, gpos); rp.pos = ix.end_key[e] - gend + 1; rp.chr = ix.end_chr[e]; }
	}
	// merge equal neighbours (elements are visited back to front for the reverse strand)
	int merged = 0;
	{
		int prev = -1;
		for (int k = 0; k < cg.n; k++) { int op = (int)(cg.e[rev ? cg.n - 1 - k : k] & 15); if (op != prev) { merged++; prev = op; } }
	}
	u32 off = KB_ALLOC(bt.cig_cursor, (u32)merged);
	rp.cig_off = off; rp.cig_len = merged;
	if ((u64)off + (u64)merged > (u64)bt.cap_cigar) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_CIGAR); rp.cig_len = 0; return; }
	int w = -1, prev = -1;
	for (int k = 0; k < cg.n; k++)
	{
		u32 e = cg.e[rev ? cg.n - 1 - k : k]; int op = (int)(e & 15);
		if (op != prev) { w++; bt.cigar[off + w] = e; prev = op; }
		else bt.cigar[off + w] += (e >> 4) << 4;
	}
}

// cigar elements and score of one candidate from its classified segments (GenMappingReport :648-723). No side effects
// besides cg / the outputs, so a candidate whose elements overflow the small local buffer is simply redone with a large one.
KB_HD void kb_assemble_cigar(const KbParams& pm, const KbBatchDev& bt, const KbSegX* sx, int n, KbCigar& cg, int* aln_out, i64* gf_out, i64* ge_out)
{
	int aln = 0;
	i64 g_first = n > 0 ? sx[0].s.gpos : 0, g_end = n > 0 ? sx[n - 1].s.gpos + sx[n - 1].s.glen - 1 : 0;
	for (int j = 0; j < n; j++)
	{
		const KbSegX& x = sx[j]; const KbSeg& sp = x.s;
		if (x.info == KB_SEG_SKIP) continue;
		if (x.info == KB_SEG_SIMPLE) { cg.push(sp.rlen, KB_OP_M); aln += sp.rlen; continue; }
		if (j == 0)   // ProcessHeadSequencePair :292 and AlignmentCandidates.cpp:669-687
		{
			int s = 0; i64 gpos = sp.gpos;
			if (x.info == KB_SEG_SOFT) cg.push(sp.rlen, KB_OP_S);
			else if (x.info == KB_SEG_QUICK) { cg.push(sp.rlen, KB_OP_M); s = (int)x.aux; }
			else
			{
				KbAlnView v = kb_view(bt, x);
				if (!kb_quality_ok(v)) cg.push(sp.rlen, KB_OP_S);
				else
				{
					int f = 0;
					if (f < v.n && (kb_view_run(v, f) & 3) == KB_RUN_D) { gpos += (int)(kb_view_run(v, f) >> 2); f++; }
					if (f < v.n && (kb_view_run(v, f) & 3) == KB_RUN_I) { cg.push((int)(kb_view_run(v, f) >> 2), KB_OP_S); f++; }
					kb_view_push(v, f, v.n, cg); s = v.ident;
				}
			}
			aln += s;
			g_first = s == 0 ? sx[1].s.gpos : gpos;
		}
		else if (j == n - 1)   // ProcessTailSequencePair :344 and AlignmentCandidates.cpp:688-706
		{
			int s = 0, glen = sp.glen;
			if (x.info == KB_SEG_SOFT) cg.push(sp.rlen, KB_OP_S);
			else if (x.info == KB_SEG_QUICK) { cg.push(sp.rlen, KB_OP_M); s = (int)x.aux; }
			else
			{
				KbAlnView v = kb_view(bt, x);
				if (!kb_quality_ok(v)) cg.push(sp.rlen, KB_OP_S);
				else
				{
					int last = v.n, clip = 0;
					if (last > 0 && (kb_view_run(v, last - 1) & 3) == KB_RUN_D) { glen -= (int)(kb_view_run(v, last - 1) >> 2); last--; }
					if (last > 0 && (kb_view_run(v, last - 1) & 3) == KB_RUN_I) { clip = (int)(kb_view_run(v, last - 1) >> 2); last--; }
					kb_view_push(v, 0, last, cg); s = v.ident;
					if (clip > 0) cg.push(clip, KB_OP_S);
				}
			}
			aln += s;
			g_end = s == 0 ? sx[j - 1].s.gpos + sx[j - 1].s.glen - 1 : sp.gpos + glen - 1;
		}
		else   // ProcessNormalSequencePair :225
		{
			if (x.info == KB_SEG_GAP) { if (sp.rlen > 0) cg.push(sp.rlen, KB_OP_I); else cg.push(sp.glen, KB_OP_D); }
			else if (x.info == KB_SEG_QUICK) { cg.push(sp.rlen, KB_OP_M); aln += (int)x.aux; }
			else { KbAlnView v = kb_view(bt, x); kb_view_push(v, 0, v.n, cg); aln += v.ident; }
		}
	}
	*aln_out = aln; *gf_out = g_first; *ge_out = g_end;
}

// phase C: GenMappingReport for one read from the stored segments and job results. cands/reports are this read's slices.
#define KB_CIG_FAST 24   // cigar elements kept in local memory before the HBM arena is used
// ar == nullptr: local memory only; returns false when a candidate's cigar does not fit (the read is then redone with an arena)
KB_HD bool kb_assemble_read(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r, KbArena* ar)
{
	KbReadRes& rd = bt.res[r];
	KbCand* cv = bt.cands + bt.cand_off[r];
	KbReport* rep = bt.reports + bt.cand_off[r];
	int ncan = bt.n_cands[r];
	bool first = pm.paired ? ((r & 1) == 0) : true;
	int rlen = (int)(bt.seq_off[r + 1] - bt.seq_off[r]);
	int score = 0, sub = 0, best = 0, best_chr = 0;
	rd.mapq = 0; rd.rep_off = bt.cand_off[r];
	if (ncan == 0)
	{
		rd.ncan = 1; rd.score = 0; rd.sub = 0; rd.best = 0;
		KbReport z; z.pos = 0; z.aln = 0; z.flag = 0; z.mate = -1; z.chr = 0; z.cig_off = 0; z.cig_len = 0; z.fwd = 1; z.pad = 0;
		rep[0] = z;
		return true;
	}
	rd.ncan = ncan;
	u32 cgl[KB_CIG_FAST];
	for (int i = 0; i < ncan; i++)
	{
		const KbCand c = cv[i];
		KbReport rp; rp.pos = 0; rp.aln = 0; rp.flag = 0; rp.mate = c.mate; rp.chr = 0; rp.cig_off = 0; rp.cig_len = 0; rp.fwd = 1; rp.pad = 0;
		if (c.score == 0) { rep[i] = rp; continue; }
		if (pm.pacbio && score > 0) { sub = score; rep[i] = rp; continue; }
		u32 ci = bt.cand_off[r] + (u32)i;
		int n = bt.cseg_n[ci];
		if (n < 0) { rep[i] = rp; continue; }      // CheckCoordinateValidity failed
		const KbSegX* sx = bt.segx + bt.cseg_off[ci];
		u64 mark = ar ? ar->used : 0;
		KbCigar cg; cg.n = 0; cg.ovf = false; cg.cap = KB_CIG_FAST; cg.e = cgl;
		i64 g_first, g_end;
		kb_assemble_cigar(pm, bt, sx, n, cg, &rp.aln, &g_first, &g_end);
		if (cg.ovf)
		{
			if (ar == nullptr) return false;
			cg.n = 0; cg.ovf = false; cg.cap = 3 * rlen + 2 * n + 64; cg.e = (u32*)ar->alloc((u64)cg.cap * 4);
			if (ar->ovf) break;
			kb_assemble_cigar(pm, bt, sx, n, cg, &rp.aln, &g_first, &g_end);
			if (cg.ovf) ar->ovf = true;
		}
		bool dead = false;
		if (!pm.pacbio && cg.n > 1)
		{
			int gp = 0; for (int k = 0; k < cg.n; k++) { int op = (int)(cg.e[k] & 15); if (op == KB_OP_I || op == KB_OP_D) gp += (int)(cg.e[k] >> 4); }   // GapPenalty :612
			rp.aln -= gp;
			if (rp.aln <= 0) { rp.aln = 0; dead = true; }
		}
		if (!dead)
		{
			if (cg.n == 0) rp.aln = 0;
			else { kb_locate_report(ix, bt, first, g_first, g_end, cg, rp); if (rp.pos <= 0) rp.aln = 0; }
			if (rp.aln > score) { best = i; best_chr = rp.chr; sub = score; score = rp.aln; }
			else if (rp.aln == score)
			{
				sub = score;
				if (!pm.multihit && ix.chr_len[rp.chr] > ix.chr_len[best_chr]) { best = i; best_chr = rp.chr; }
			}
		}
		rep[i] = rp;
		if (ar) { ar->used = mark; if (ar->ovf) break; }
	}
	rd.score = score; rd.sub = sub; rd.best = best;
	return true;
}

#endif
