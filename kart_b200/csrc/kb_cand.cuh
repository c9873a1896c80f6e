// Candidate construction, pruning and mate pairing (one thread per read pair / per single read).
// Replaces: sort by (PosDiff,rPos) src/AlignmentCandidates.cpp:77, GenerateAlignmentCandidateForIlluminaSeq :82,
// sort by (gPos,rPos) :166, GenerateAlignmentCandidateForPacBioSeq :171, RemoveRedundantCandidates src/Mapping.cpp:317,
// CheckPairedAlignmentCandidates :348, RemoveUnMatedAlignmentCandidates :402.
#ifndef KB_CAND_CUH
#define KB_CAND_CUH
#include "kb_fm.cuh"

// ChrLocMap.lower_bound(pos): index of the first end key >= pos, or n_ends. The keys are few (two per sequence) but a binary
// search over them is a chain of dependent global loads per call, and candidate construction, segment validation and
// coordinate output each call it per candidate (ncu r17, C3: 22 % of k_cand_pair's and 16 % of k_segments' stall samples). A small
// table of "first key at or behind the start of this 2^end_shift-base bucket" leaves a range of 0..2 keys for typical genomes.
KB_HD int kb_chr_lookup(const KbIndexDev& ix, i64 pos)
{
	int lo = 0, hi = ix.n_ends;
	if (ix.end_tab != nullptr)
	{
		if (pos <= 0) return 0;
		if (pos >= ix.G2) return ix.n_ends;
		const i64 b = pos >> ix.end_shift;
		lo = (int)KB_LDG(ix.end_tab + b); hi = (int)KB_LDG(ix.end_tab + b + 1);
	}
	while (lo < hi) { int mid = (lo + hi) >> 1; if (ix.end_key[mid] < pos) lo = mid + 1; else hi = mid; }
	return lo;
}

KB_HD bool kb_less_diff(const KbSeg& a, const KbSeg& b)
{
	i64 da = a.gpos - a.rpos, db = b.gpos - b.rpos;
	return da == db ? a.rpos < b.rpos : da < db;
}
KB_HD bool kb_less_gpos(const KbSeg& a, const KbSeg& b) { return a.gpos == b.gpos ? a.rpos < b.rpos : a.gpos < b.gpos; }

// In-place sorts of a short list. Both orders are total on distinct seeds, so any correct sort reproduces std::sort.
// Shell sort keeps the rare long lists (repeats, pacbio) out of quadratic time.
template <bool BY_GPOS>
KB_HD void kb_sort_segs(KbSeg* v, int n)
{
	int gap = 1; while (gap < n / 3) gap = gap * 3 + 1;
	for (; gap >= 1; gap /= 3)
		for (int i = gap; i < n; i++)
		{
			KbSeg t = v[i]; int j = i;
			while (j >= gap && (BY_GPOS ? kb_less_gpos(t, v[j - gap]) : kb_less_diff(t, v[j - gap]))) { v[j] = v[j - gap]; j -= gap; }
			v[j] = t;
		}
}

// Illumina candidates: runs of (PosDiff,rPos)-sorted seeds whose neighbours differ by <= MaxGaps on the diagonal
// and stay on the first seed's chromosome; kept when the summed seed length beats a ratcheting threshold.
// WIDE: the seed list is long and in HBM (k_cand_heavy_finish): the scan fetches four seeds per round of loads.
template <bool WIDE>
KB_HD int kb_cands_illumina(const KbIndexDev& ix, const KbParams& pm, int rlen, KbSeg* sv, int n, u32 seg_base, KbCand* out, int cap)
{
	int thr = (int)(rlen * 0.2); if (thr > 50) thr = 50;
	int nc = 0, i = 0;
	while (i < n && sv[i].gpos - sv[i].rpos < 0) i++;
	while (i < n)
	{
		int score = sv[i].rlen, k;
		i64 bound = ix.end_key[kb_chr_lookup(ix, sv[i].gpos)];
		// four seeds per round of loads: a heavy item (hundreds of seeds, one thread: k_cand_heavy_finish) otherwise pays one memory round trip
		// per seed, the break test standing between consecutive loads (ncu r29: 30 % of that kernel's samples on this line, at 6 % warps active)
		i64 dprev = sv[i].gpos - sv[i].rpos; bool open = true;
		if (!WIDE)
		{
			for (k = i + 1; k < n; k++)
			{
				const i64 d = sv[k].gpos - sv[k].rpos;
				if (sv[k].gpos > bound || d - dprev > pm.max_gaps) break;
				score += sv[k].rlen; dprev = d;
			}
		}
		else for (k = i + 1; k < n && open;)
		{
			const int m = n - k < 4 ? n - k : 4;
			i64 g[4]; int rp[4], rl[4];
			for (int u = 0; u < 4; u++) if (u < m) { g[u] = sv[k + u].gpos; rp[u] = sv[k + u].rpos; rl[u] = sv[k + u].rlen; }
			for (int u = 0; u < 4; u++)
			{
				if (u >= m) break;
				const i64 d = g[u] - rp[u];
				if (g[u] > bound || d - dprev > pm.max_gaps) { open = false; break; }
				score += rl[u]; dprev = d; k++;
			}
		}
		if (score > thr)
		{
			if (score - 50 > thr) thr = score - 50;
			i64 d = sv[i].gpos - sv[i].rpos;
			if (nc < cap) { KbCand c; c.score = score; c.mate = -1; c.diff = d < 0 ? 0 : d; c.seg_start = seg_base + (u32)i; c.nseg = k - i; out[nc] = c; }
			nc++;
			kb_sort_segs<true>(sv + i, k - i);
		}
		i = k;
	}
	return nc;
}

// PacBio candidates: greedy chaining over (gPos,rPos)-sorted seeds. Chains are not contiguous in the seed list, so
// each kept chain is copied to freshly allocated seed storage. `taken` is per-thread scratch (n bytes).
KB_HD int kb_cands_pacbio(const KbBatchDev& bt, const KbSeg* sv, int n, u8* taken, KbSeg* tmp, KbCand* out, int cap)
{
	int nc = 0, thr = 0, i = 0;
	for (int t = 0; t < n; t++) taken[t] = 0;
	while (i < n && sv[i].gpos - sv[i].rpos < 0) i++;
	for (; i < n; i++)
	{
		if (taken[i]) continue;
		int score = sv[i].rlen, j = i, m = 0;
		taken[i] = 1; tmp[m++] = sv[i];
		for (int k = i + 1; k < n; k++)
		{
			if (taken[k]) continue;
			i64 d = (sv[k].gpos - sv[k].rpos) - (sv[j].gpos - sv[j].rpos); if (d < 0) d = -d;
			if (d < 300)
			{
				if (sv[k].rpos > sv[j].rpos) { score += sv[k].rlen; tmp[m++] = sv[k]; taken[k] = 1; j = k; }
			}
			else if (sv[k].gpos - sv[j].gpos > 1000) break;
		}
		if (score >= thr)
		{
			thr = score;
			u32 off = KB_ALLOC(&bt.counters[0], (u32)m);
			if ((u64)off + (u64)m > (u64)bt.cap_segs) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEEDS); return nc; }
			for (int t = 0; t < m; t++) bt.segs[off + t] = tmp[t];
			i64 d = sv[i].gpos - sv[i].rpos;
			if (nc < cap) { KbCand c; c.score = score; c.mate = -1; c.diff = d < 0 ? 0 : d; c.seg_start = off; c.nseg = m; out[nc] = c; }
			nc++;
		}
	}
	return nc;
}

KB_HD void kb_prune(const KbParams& pm, KbCand* v, int n)   // RemoveRedundantCandidates
{
	if (n <= 1) return;
	int s1 = 0, s2 = 0;
	for (int i = 0; i < n; i++) if (v[i].score > s2) { if (v[i].score >= s1) { s2 = s1; s1 = v[i].score; } else s2 = v[i].score; }
	int thr = (pm.pacbio || s1 == s2 || s1 - s2 > 20) ? s1 : s2;
	for (int i = 0; i < n; i++) if (v[i].score < thr) v[i].score = 0;
}

// CheckPairedAlignmentCandidates. Also narrows [*lo,*hi], the interval of EstDistance values for which every
// `dist < Est` comparison made here has the same outcome (used by the host to keep the per-chunk recurrence exact).
KB_HD bool kb_pair(const KbParams& pm, i64 est, KbCand* a, int n1, KbCand* b, int n2, i32* lo, i32* hi)
{
	bool any = false;
	if (n1 * n2 > 1000) { kb_prune(pm, a, n1); kb_prune(pm, b, n2); }
	// Both lists come out of a (PosDiff,rPos)-sorted seed scan, so their PosDiff ascends: the mates of a[i] that the reference's
	// inner loop does anything with (PosDiff >= a[i]'s, distance < Est) are one window of b, followed by at most one candidate that
	// narrows *hi; every later one is further away still. The window's start only moves forward with i. A pair inside a repeat
	// family has dozens of candidates on either side, all but a few of them a genome away from each other: n1 + n2 steps instead of
	// n1 x n2 (ncu r21, C3: the quadratic loop was a quarter of k_cand_heavy). The order is checked, not assumed.
	bool ascending = n1 * n2 > 16;
	if (ascending)   // no early exit: the loads of consecutive candidates then overlap instead of waiting for each other's test
	{
		bool asc = true;
		for (int i = 1; i < n1; i++) asc &= !(a[i].diff < a[i - 1].diff);
		for (int j = 1; j < n2; j++) asc &= !(b[j].diff < b[j - 1].diff);
		ascending = asc;
	}
	int j0 = 0;
	for (int i = 0; i < n1; i++)
	{
		if (a[i].score == 0) continue;
		int best = -1, s = 0;
		if (ascending) while (j0 < n2 && b[j0].diff < a[i].diff) j0++;
		for (int j = j0; j < n2; j++)
		{
			if (b[j].score == 0 || b[j].diff < a[i].diff) continue;
			i64 dist = b[j].diff - a[i].diff;
			if (dist < est)
			{
				if (dist + 1 > *lo) *lo = (i32)(dist + 1);
				if (b[j].score > s) { best = j; s = b[j].score; } else if (b[j].score == s) best = -1;
			}
			else { if (dist < *hi) *hi = (i32)dist; if (ascending) break; }
		}
		if (s > 0 && best != -1)
		{
			int j = best;
			if (b[j].mate == -1) { any = true; a[i].mate = j; b[j].mate = i; }
			else if (a[i].score > a[b[j].mate].score) { a[b[j].mate].mate = -1; a[i].mate = j; b[j].mate = i; }
		}
	}
	return any;
}

KB_HD void kb_keep_mated(KbCand* a, int n1, KbCand* b, int n2)   // RemoveUnMatedAlignmentCandidates
{
	for (int i = 0; i < n1; i++)
	{
		if (a[i].mate == -1) a[i].score = 0;
		else { int j = a[i].mate; int s = a[i].score + b[j].score; a[i].score = s; b[j].score = s; }
	}
	for (int j = 0; j < n2; j++) if (b[j].mate == -1) b[j].score = 0;
}

#endif
