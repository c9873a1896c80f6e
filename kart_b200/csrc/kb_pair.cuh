// Paired-end rescue and final per-pair scoring: flags, MAPQ, insert-size statistic.
// Replaces: RescueUnpairedAlignment src/AlignmentRescue.cpp:73 (+ IdnetifyRescueCandidate :26, DetermineAnchorThreshold :14),
// CheckPairedFinalAlignments src/Mapping.cpp:429, SetPairedAlignmentFlag :73, SetSingleAlignmentFlag :49, EvaluateMAPQ :160,
// and the (iPaired, iDistance) bookkeeping of OutputPairedAlignments :206-213.
#ifndef KB_PAIR_CUH
#define KB_PAIR_CUH
#include "kb_align.cuh"

KB_HD int kb_top_score(const KbCand* v, int n) { int s = 0; for (int i = 0; i < n; i++) if (v[i].score > s) s = v[i].score; return s; }

// Best diagonal cluster of the exact matches found in one reference window (IdnetifyRescueCandidate).
// v is in (PosDiff,rPos) order with window-relative gPos. On success the cluster's seeds are written to the global
// seed arena, (gPos,rPos)-sorted, and *out describes the candidate.
KB_HD int kb_rescue_cluster(const KbParams& pm, const KbBatchDev& bt, i64 left, KbSeg* v, int n, KbCand* out)
{
	int best_s = 0, best_i = 0, best_n = 0;
	for (int i = 0; i < n;)
	{
		int s = v[i].rlen, j; i64 di = v[i].gpos - v[i].rpos;
		for (j = i + 1; j < n; j++) { if ((v[j].gpos - v[j].rpos) - di < pm.max_gaps) s += v[j].rlen; else break; }
		if (s > best_s) { best_s = s; best_i = i; best_n = j - i; }
		i = j;
	}
	out->score = best_s; out->mate = -1; out->nseg = 0; out->seg_start = 0; out->diff = 0;
	if (best_s == 0) return 0;
	out->diff = (v[best_i].gpos - v[best_i].rpos) + left;
	u32 off = KB_ATOMIC_ADD(&bt.counters[0], (u32)best_n);
	if ((u64)off + (u64)best_n > (u64)bt.cap_segs) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SEEDS); out->score = 0; return 0; }
	kb_sort_segs<true>(v + best_i, best_n);
	for (int k = 0; k < best_n; k++) { KbSeg s = v[best_i + k]; s.gpos += left; bt.segs[off + k] = s; }
	out->seg_start = off; out->nseg = best_n;
	return best_s;
}

// Search `mate` (length ml) in the reference window [left, left+slen) with 8-mers; exact runs >= 10 (AlignmentRescue.cpp:119-122).
KB_HD int kb_rescue_window(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, KbArena& ar, const u32* wm, int ml, i64 left, int slen, KbCand* out)
{
	u64 mark = ar.used; int score = 0;
	u8* win = (u8*)ar.alloc((u64)slen);
	u32* ww = (u32*)ar.alloc((u64)slen * 4);
	u64 worst = (u64)((ml < slen ? ml : slen) / 9 + 2) * (u64)(ml + slen);
	u64 room = ar.cap > ar.used ? (ar.cap - ar.used) / (2 * sizeof(KbSeg)) : 0;
	int cap = (int)(worst < room ? worst : room);
	KbSeg* pairs = (KbSeg*)ar.alloc((u64)(cap > 0 ? cap : 1) * sizeof(KbSeg));
	if (!ar.ovf)
	{
		for (int i = 0; i < slen; i++) win[i] = kb_code_char(kb_ref_code(ix, left + i));
		kb_kmer_ids(slen, win, ww);
		bool povf = false;
		int np = kb_kmer_pairs(wm, ml, ww, slen, slen, 10, pairs, cap, &povf);
		if (povf) ar.ovf = true;
		else score = kb_rescue_cluster(pm, bt, left, pairs, np, out);
	}
	ar.used = mark;
	return score;
}

// RescueUnpairedAlignment for one pair. a/b: candidate slices of read 1 / read 2 with capacities ca/cb.
KB_HD bool kb_rescue_pair(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, KbArena& ar, int est,
                          const u8* s1, int l1, const u8* s2, int l2, KbCand* a, int* pn1, int ca, KbCand* b, int* pn2, int cb, bool* attempted)
{
	int n1 = *pn1, n2 = *pn2;
	int sc1 = kb_top_score(a, n1), sc2 = kb_top_score(b, n2), strategy; bool mated = false;
	*attempted = false;
	if (sc1 == 0 && sc2 == 0) return false;
	if (sc1 < (int)(l1 * 0.1) && sc2 < (int)(l2 * 0.1)) strategy = 4;
	else if (sc1 > sc2 && sc1 - sc2 > 50) strategy = 1;
	else if (sc2 > sc1 && sc2 - sc1 > 50) strategy = 2;
	else strategy = 3;
	if (strategy == 4) return false;
	*attempted = true;
	if (est > pm.max_insert) est = pm.max_insert;
	u64 mark = ar.used;
	int lm = l1 > l2 ? l1 : l2;
	u32* wm = (u32*)ar.alloc((u64)lm * 4);
	if (ar.ovf) return false;
	if (strategy == 1 || strategy == 3)
	{
		int thr = sc1 - 30 < 50 ? 50 : sc1 - 30;
		kb_kmer_ids(l2, s2, wm);
		for (int j = n2, i = 0; i < n1 && !ar.ovf; i++)
		{
			if (a[i].score < thr) continue;
			i64 left = a[i].diff, right = a[i].diff + est + l2;
			int e = kb_chr_lookup(ix, left); if (e >= ix.n_ends) continue;
			int cid = ix.end_chr[e];
			if (right < ix.G && right > ix.chr_fwd[cid]) right = ix.chr_fwd[cid] - 1;
			else if (right >= ix.G && right > ix.chr_rev[cid]) right = ix.chr_rev[cid] - 1;
			int slen = (int)(right - left); if (slen < l2) continue;
			KbCand c;
			if (kb_rescue_window(ix, pm, bt, ar, wm, l2, left, slen, &c) > sc2 && j < cb) { mated = true; c.mate = i; a[i].mate = j; b[j++] = c; *pn2 = j; }
		}
	}
	if (strategy == 2 || strategy == 3)
	{
		int thr = sc2 - 30 < 50 ? 50 : sc2 - 30;
		kb_kmer_ids(l1, s1, wm);
		for (int i = n1, j = 0; j < n2 && !ar.ovf; j++)
		{
			if (b[j].score < thr) continue;
			i64 left = b[j].diff - est, right = b[j].diff + l2;
			int e = kb_chr_lookup(ix, right); if (e >= ix.n_ends) continue;
			int cid = ix.end_chr[e];
			if (left < ix.G && left < ix.chr_fwd[cid] - ix.chr_len[cid]) left = ix.chr_fwd[cid] - ix.chr_len[cid] + 1;
			else if (right >= ix.G && left < ix.chr_rev[cid] - ix.chr_len[cid]) left = ix.chr_rev[cid] - ix.chr_len[cid] + 1;
			int slen = (int)(right - left); if (slen < l1) continue;
			KbCand c;
			if (kb_rescue_window(ix, pm, bt, ar, wm, l1, left, slen, &c) > sc1 && i < ca) { mated = true; c.mate = j; b[j].mate = i; a[i++] = c; *pn1 = i; }
		}
	}
	ar.used = mark;
	return mated;
}

// ---- final scoring -------------------------------------------------------------------------------
KB_HD void kb_settle_pair(const KbParams& pm, KbReadRes& r1, KbReport* p1, KbReadRes& r2, KbReport* p2)   // CheckPairedFinalAlignments
{
	bool mated = p1[r1.best].mate == r2.best;
	if (!pm.multihit && mated) return;
	if (!mated && r1.score > 0 && r2.score > 0)
	{
		int s = 0;
		for (int i = 0; i < r1.ncan; i++)
		{
			int j = p1[i].mate;
			if (p1[i].aln > 0 && j != -1 && p2[j].aln > 0)
			{
				mated = true;
				if (s < p1[i].aln + p2[j].aln) { s = p1[i].aln + p2[j].aln; r1.best = i; r1.score = p1[i].aln; r2.best = j; r2.score = p2[j].aln; }
			}
		}
	}
	if (mated)
	{
		for (int i = 0; i < r1.ncan; i++)
		{
			int j = p1[i].mate;
			if (p1[i].aln != r1.score || (j != -1 && p2[j].aln != r2.score)) { p1[i].aln = 0; p1[i].mate = -1; }
		}
	}
	else
	{
		for (int i = 0; i < r1.ncan; i++) { p1[i].mate = -1; if (p1[i].aln > 0 && p1[i].aln != r1.score) p1[i].aln = 0; }
		for (int j = 0; j < r2.ncan; j++) { p2[j].mate = -1; if (p2[j].aln > 0 && p2[j].aln != r2.score) p2[j].aln = 0; }
	}
}

KB_HD void kb_flag_single(KbReadRes& r, KbReport* p)   // SetSingleAlignmentFlag
{
	if (r.score > r.sub) p[r.best].flag = p[r.best].fwd ? 0 : 0x10;
	else if (r.score > 0) { for (int i = 0; i < r.ncan; i++) if (p[i].aln > 0) p[i].flag = p[i].fwd ? 0 : 0x10; }
	else p[0].flag = 0x4;
}

KB_HD void kb_flag_half(KbReadRes& me, KbReport* pm_, const KbReadRes& other, const KbReport* po, int base)   // SetPairedAlignmentFlag :96-156
{
	if (me.score > me.sub)
	{
		KbReport& a = pm_[me.best]; a.flag = base | (a.fwd ? 0x20 : 0x10);
		if (a.mate != -1 && po[a.mate].aln > 0) a.flag |= 0x2; else a.flag |= 0x8;
	}
	else if (me.score > 0)
	{
		for (int i = 0; i < me.ncan; i++)
		{
			KbReport& a = pm_[i]; if (a.aln <= 0) continue;
			a.flag = base | (a.fwd ? 0x20 : 0x10);
			if (a.mate != -1 && po[a.mate].aln > 0) a.flag |= 0x2; else a.flag |= 0x8;
		}
	}
	else
	{
		pm_[0].flag = base | 0x4;
		if (other.score == 0) pm_[0].flag |= 0x8; else pm_[0].flag |= (po[other.best].fwd ? 0x10 : 0x20);
	}
}

KB_HD void kb_flag_pair(KbReadRes& r1, KbReport* p1, KbReadRes& r2, KbReport* p2)   // SetPairedAlignmentFlag
{
	if (r1.score > r1.sub && r2.score > r2.sub)
	{
		KbReport& a = p1[r1.best]; KbReport& b = p2[r2.best];
		a.flag = 0x41; b.flag = 0x81;
		if (r2.best == a.mate) { a.flag |= 0x2; b.flag |= 0x2; }
		a.flag |= a.fwd ? 0x20 : 0x10; b.flag |= b.fwd ? 0x20 : 0x10;
	}
	else { kb_flag_half(r1, p1, r2, p2, 0x41); kb_flag_half(r2, p2, r1, p1, 0x81); }
}

// EvaluateMAPQ. The Illumina branch mixes float, double and libm log(); it is evaluated on the host once per (score, diff)
// with the reference expression and looked up here. Scores beyond the table fall back to device arithmetic.
KB_HD void kb_mapq(const KbIndexDev& ix, const KbParams& pm, KbReadRes& r, int rlen)
{
	if (r.score == 0 || r.score == r.sub) { r.mapq = 0; return; }
	if (pm.pacbio)
	{
		float scale = (float)(85.0 * (int)(ceil((double)(rlen / 100) + 0.5)));
		if (scale > 2000) scale = 2000;
		r.mapq = (int)(60 * (r.score / scale));
	}
	else if (r.sub == 0 || r.score - r.sub > 5) r.mapq = 60;
	else if (r.score < ix.mapq_lut_scores) r.mapq = ix.mapq_lut[r.score * 5 + (r.score - r.sub - 1)];
	else r.mapq = (int)(30 * (1 - (float)(r.score - r.sub) / r.score) * log((double)r.score) + 0.4999);
	if (r.mapq > 60) r.mapq = 60;
}

KB_HD void kb_finalize_pair(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int p)
{
	int ra = 2 * p, rb = ra + 1;
	KbReadRes& r1 = bt.res[ra]; KbReadRes& r2 = bt.res[rb];
	KbReport* p1 = bt.reports + r1.rep_off; KbReport* p2 = bt.reports + r2.rep_off;
	int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
	kb_settle_pair(pm, r1, p1, r2, p2);
	kb_flag_pair(r1, p1, r2, p2);
	kb_mapq(ix, pm, r1, l1); kb_mapq(ix, pm, r2, l2);
	KbPairStat& st = bt.pstat[p]; st.counted = 0; st.absdist = 0;
	if (r1.score > 0)
	{
		const KbReport& x = p1[r1.best]; int j = x.mate;
		if (x.aln > 0 && j != -1 && p2[j].aln > 0)
		{
			int dist = (int)(p2[j].pos - x.pos + (x.fwd ? l2 : 0 - l1));
			st.counted = 1; st.absdist = dist < 0 ? -dist : dist;
		}
	}
}

KB_HD void kb_finalize_single(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r)
{
	KbReadRes& rd = bt.res[r];
	kb_flag_single(rd, bt.reports + rd.rep_off);
	kb_mapq(ix, pm, rd, (int)(bt.seq_off[r + 1] - bt.seq_off[r]));
}

#endif
