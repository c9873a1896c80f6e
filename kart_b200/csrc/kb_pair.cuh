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

// ---- block-cooperative window search -----------------------------------------------------------------
// One thread block per reference window of a pair that failed to pair (RescueUnpairedAlignment; the windows of a job are
// enumerated by kb_rescue_plan below). The threads share the work: look every window 8-mer up in a small index of the mate's
// 8-mers and extend the run starts. The record lives in shared memory; the phases are plain functions of (record, tid, nth)
// that the host-emulation build can replay.
struct KbRescueJob   // one reference window of one rescue job
{
	i32 p, ra, rb, l1, l2;           // the pair, its two reads and their lengths
	i32 side, idx;                   // 0: read 2 is searched around candidate idx of read 1 ; 1: the other way round
	i32 done, ovf, clean, mclean, slen, cap_pairs, ml, reindex, hmask;
	u32 npairs;
	i64 left;
	u32* wm; u8* win; u32* ww; KbSeg* pairs;
	u32* hkey; i32* hhead; i32* hnext;   // open-addressing index of the mate's 8-mers: id+1 per slot, chain of positions per slot
	const u8* mate; const KbPk* mate_pk;
};

KB_HD KbArena kb_job_arena(const KbBatchDev& bt, int block, int nth)   // the block's share of the per-thread scratch
{
	KbArena ar; u64 per = bt.scratch_per_thread * (u64)nth;
	ar.base = bt.scratch + (u64)block * per; ar.used = 0; ar.cap = per; ar.ovf = false;
	return ar;
}
// 8-mer id at window position g / mate position r. Inside the text (and for a mate of pure bases) the id is the 16 bits of the
// packed sequence at that position, so neither the window characters nor the id arrays are materialised.
KB_HD u32 kb_rj_wid(const KbIndexDev& ix, const KbRescueJob* j, int g)
{
	if (!j->clean) return j->ww[g];
	if (g + 8 > j->slen) return KB_NOKMER;
	u32 inv; return (u32)(kb_ref_win(ix, j->left + g, &inv) >> 48);
}
KB_HD u32 kb_rj_mid(const KbRescueJob* j, int r)
{
	if (!j->mclean) return j->wm[r];
	if (r + 8 > j->ml) return KB_NOKMER;
	return (u32)(kb_read_win(j->mate_pk, r).code >> 48);
}

// all threads: reference characters of the window
KB_HD void kb_rj_window(const KbIndexDev& ix, KbRescueJob* j, int tid, int nth)
{
	if (j->clean) return;
	for (int i = tid; i < j->slen; i += nth) j->win[i] = kb_code_char(kb_ref_code(ix, j->left + i));
}

// all threads: 8-mer ids of the window. A window inside the text is pure ACGT, where the reference's rolling ids equal the
// 16-bit value of the 8 characters at every position; otherwise thread 0 replays the literal scan.
KB_HD void kb_rj_ids(KbRescueJob* j, int tid, int nth)
{
	if (!j->clean && tid == 0) kb_kmer_ids(j->slen, j->win, j->ww);
}

// all threads, after a side switch: index the mate's 8-mers (two barrier-separated phases: clear, insert)
KB_HD u32 kb_rj_slot(u32 id, int mask) { return (id * 40503u + (id >> 7)) & (u32)mask; }
KB_HD void kb_rj_index_clear(const KbBatchDev& bt, KbRescueJob* j, int tid, int nth)
{
	if (!j->reindex) return;
	for (int s = tid; s <= j->hmask; s += nth) { j->hkey[s] = 0; j->hhead[s] = -1; }
	// 8-mer ids of the mate (CreateKmerVecFromReadSeq): a mate of pure bases has id(p) = its 16 packed bits at p; anything
	// else replays the literal scan on one lane
	const KbPk* rd = kb_pk_read(bt, j->side == 0 ? j->rb : j->ra);
	bool dirty = false;
	for (int w = 0; 32 * w < j->ml; w++) { u32 n4 = kb_load_pk(rd + w).n4; int rem = j->ml - 32 * w; if (rem < 32) n4 &= ~(~0u >> rem); if (n4) dirty = true; }
	if (tid == 0) { j->mclean = dirty ? 0 : 1; j->mate_pk = rd; if (dirty) kb_kmer_ids(j->ml, j->mate, j->wm); }
}
KB_HD void kb_rj_index_fill(KbRescueJob* j, int tid, int nth)
{
	if (!j->reindex) return;
	for (int r = tid; r < j->ml; r += nth)
	{
		u32 id = kb_rj_mid(j, r); if (id == KB_NOKMER) continue;
		u32 s = kb_rj_slot(id, j->hmask);
		while (true)
		{
			u32 old = KB_ATOMIC_CAS(&j->hkey[s], 0u, id + 1u);
			if (old == 0u || old == id + 1u) break;
			s = (s + 1u) & (u32)j->hmask;
		}
		j->hnext[r] = KB_ATOMIC_EXCH(&j->hhead[s], (i32)r);
	}
}

// all threads: exact runs of >= 10 bases between the mate and the window, appended in arbitrary order. Same set as
// IdentifyCommonKmers(MaxShift = slen) + GenerateSimplePairsFromCommonKmers(10): every window position looks its 8-mer up
// in the mate's index; a match (r,g) whose predecessor (r-1,g-1) is no match starts a run, which is then extended.
// (|g - r| < slen holds for every pair since r < ml <= slen and g < slen.)
KB_HD void kb_rj_pairs(const KbIndexDev& ix, KbRescueJob* j, int tid, int nth)
{
	const int ml = j->ml, sl = j->slen;
	for (int g = tid; g < sl; g += nth)
	{
		u32 id = kb_rj_wid(ix, j, g); if (id == KB_NOKMER) continue;
		u32 s = kb_rj_slot(id, j->hmask), key;
		while ((key = j->hkey[s]) != 0u && key != id + 1u) s = (s + 1u) & (u32)j->hmask;
		if (key == 0u) continue;
		for (int r = j->hhead[s]; r >= 0; r = j->hnext[r])
		{
			if (r > 0 && g > 0) { u32 a = kb_rj_mid(j, r - 1); if (a != KB_NOKMER && a == kb_rj_wid(ix, j, g - 1)) continue; }   // not the start of its run
			int run = 1;
			while (r + run < ml && g + run < sl) { u32 a = kb_rj_mid(j, r + run); if (a == KB_NOKMER || a != kb_rj_wid(ix, j, g + run)) break; run++; }
			int l = 8 + run - 1;
			if (l < 10) continue;
			u32 slot = KB_ATOMIC_ADD(&j->npairs, 1u);
			if ((int)slot < j->cap_pairs) { KbSeg sg; sg.simple = 1; sg.rpos = r; sg.gpos = (i64)g; sg.rlen = sg.glen = l; j->pairs[slot] = sg; }
			else j->ovf = 1;
		}
	}
}

// ---- warp-per-window fast path ---------------------------------------------------------------------------
// Nearly every rescue window lies inside the text (pure ACGT) and faces a mate of pure bases. Then the 8-mer ids on both sides
// are 16 bits of the packed sequences, "the 8-mers at (r,g) and (r-1,g-1) both match" is "base r-1 equals base g-1", and a run of
// matching 8-mers is an exact base match: the window's runs of >= 10 bases are the left-maximal exact matches between mate and
// window, each found where a window 8-mer hits the mate's 8-mer index and measured by XOR + count-leading-zeros on 32-base
// words. One warp does a window out of shared memory: the mate's packed words, the window's packed words (one kb_ref_win per 32
// positions instead of one per position), a 512-slot open-addressing index of the mate's 8-mers and the run list. ncu r17 (C3, 3.1 Gbp:
// 244 k windows per million reads) had the block-per-window version at 44 k warp instructions per window, a third of them
// probing the index in the HBM arena; this one needs about a thousand. Windows it cannot take (outside the text, a mate with
// other characters or longer than 255, a window of more than 2048 positions, more than KB_RF_PAIRS runs) go to the slow list and
// through the block-per-window phases above unchanged.
#define KB_RF_SLOTS 512
#define KB_RF_WORDS 66      // window words: 2048 positions + the word the last 8-mers reach into + 1
#define KB_RF_PAIRS 24
#define KB_RF_FILT 16384
#define KB_RF_CAND 256
KB_HD u32 kb_rf_fold(u32 id) { return (id ^ (id >> 2)) & (u32)(KB_RF_FILT - 1); }
struct KbRescueFast
{
	u32 hkey[KB_RF_SLOTS]; u32 hhead[KB_RF_SLOTS];
	u32 filt[KB_RF_FILT / 32];   // one bit per 14-bit fold of an 8-mer id: set where the mate has such an 8-mer (99 % of window positions stop here)
	u64 wcode[KB_RF_WORDS]; u64 mcode[10];
	KbSeg pairs[KB_RF_PAIRS];
	u8 hnext[256];
	unsigned short cand[KB_RF_CAND];   // window positions that passed the filter (kb_rf_scan -> kb_rf_pairs)
	u32 task, npairs, ncand, cand_cap; i32 ok, dirty, ml, slen, mate_read, stride; i64 left;
	i32 have_mate, same;   // the mate whose index is in place (-1: none) ; this task faces the same mate: index, filter and packed mate are kept
};
KB_HD u64 kb_rf_bits(const u64* w, int p)   // 32 bases starting at base p of a packed word array (one spare word behind the data)
{
	const int s = (p & 31) * 2; const u64 a = w[p >> 5];
	return s ? (a << s) | (w[(p >> 5) + 1] >> (64 - s)) : a;
}
// lane 0: the task's window and mate; decides whether the fast path applies
KB_HD void kb_rf_begin(const KbIndexDev& ix, const KbBatchDev& bt, KbRescueFast& w, u32 task)
{
	const KbRTask t = bt.rtasks[task];
	const int p = bt.rescue_list[t.job], ra = 2 * p, rb = ra + 1;
	w.task = task; w.npairs = 0; w.ncand = 0; w.cand_cap = (u32)(bt.rf_cand < KB_RF_CAND ? (bt.rf_cand > 0 ? bt.rf_cand : 0) : KB_RF_CAND); w.stride = bt.rf_stride == 1 ? 1 : 3;
	w.mate_read = t.side == 0 ? rb : ra;
	w.left = t.left; w.slen = t.slen;
	// a job's windows of one side follow each other and face the same mate: its 8-mer index is built once per run of such tasks
	w.same = (w.have_mate == w.mate_read) ? 1 : 0;
	if (!w.same) { w.ml = (int)(bt.seq_off[w.mate_read + 1] - bt.seq_off[w.mate_read]); w.dirty = 0; }
	w.ok = (t.left >= 0 && t.left + (i64)t.slen <= ix.G2 && t.slen <= 2048 && w.ml >= 8 && w.ml <= 255) ? 1 : 0;
	if (!w.same) w.have_mate = (w.ok && bt.rf_reuse) ? w.mate_read : -1;   // set before the lanes fill it: a task that is not ok leaves the old index untouched but unusable
}
// all lanes: clear the index, fetch the packed mate (flagging characters that are no bases) and the packed window
KB_HD void kb_rf_load(const KbIndexDev& ix, const KbBatchDev& bt, KbRescueFast& w, int lane)
{
	if (!w.ok) return;
	if (!w.same)
	{
		for (int s = lane; s < KB_RF_SLOTS; s += 32) { w.hkey[s] = 0; w.hhead[s] = 0xFFFFFFFFu; w.filt[s] = 0; }
		const KbPk* rd = kb_pk_read(bt, w.mate_read);
		const int mw = (w.ml + 31) >> 5;
		for (int k = lane; k < 10; k += 32)
		{
			u64 code = 0;
			if (k < mw) { const KbPk v = kb_load_pk(rd + k); code = v.code; u32 n4 = v.n4; const int rem = w.ml - 32 * k; if (rem < 32) n4 &= ~(~0u >> rem); if (n4) w.dirty = 1; }
			w.mcode[k] = code;
		}
	}
	const int ww = ((w.slen + 31) >> 5) + 1;
	for (int k = lane; k < ww; k += 32) { u32 inv; w.wcode[k] = kb_ref_win(ix, w.left + 32 * (i64)k, &inv); }
}
// all lanes: index the mate's 8-mers
KB_HD void kb_rf_fill(KbRescueFast& w, int lane)
{
	if (!w.ok || w.dirty || w.same) return;
	for (int r = lane; r + 8 <= w.ml; r += 32)
	{
		const u32 id = (u32)(kb_rf_bits(w.mcode, r) >> 48);
		u32 s = kb_rj_slot(id, KB_RF_SLOTS - 1);
		while (true)
		{
			const u32 old = KB_ATOMIC_CAS(&w.hkey[s], 0u, id + 1u);
			if (old == 0u || old == id + 1u) break;
			s = (s + 1u) & (u32)(KB_RF_SLOTS - 1);
		}
		w.hnext[r] = (u8)KB_ATOMIC_EXCH(&w.hhead[s], (u32)r);
		const u32 h = kb_rf_fold(id);
		KB_ATOMIC_OR(&w.filt[h >> 5], 1u << (h & 31));
	}
}
// one lane: window position g passed the filter: look its 8-mer up in the mate's index and measure the runs that start here.
// SAMPLED (every third window position is scanned, see kb_rf_scan): the 8-mer at (r, g) may lie up to two bases inside its run, so
// the run is followed to the left first; three matching bases to the left mean that the sampled position g - 3 is inside the same
// run and reports it.
template <bool SAMPLED>
KB_HD void kb_rf_probe(KbRescueFast& w, int g0, u32 id)
{
	const u64 M5 = 0x5555555555555555ull;
	const int ml = w.ml, sl = w.slen;
	u32 s = kb_rj_slot(id, KB_RF_SLOTS - 1), key;
	while ((key = w.hkey[s]) != 0u && key != id + 1u) s = (s + 1u) & (u32)(KB_RF_SLOTS - 1);
	if (key == 0u) return;
	for (u32 r0 = w.hhead[s]; r0 < 255u; r0 = w.hnext[r0])
	{
		int r = (int)r0, g = g0, back = 0;
		while (r > 0 && g > 0 && ((w.mcode[(r - 1) >> 5] >> (62 - 2 * ((r - 1) & 31))) & 3ull) == ((w.wcode[(g - 1) >> 5] >> (62 - 2 * ((g - 1) & 31))) & 3ull))
		{
			if (!SAMPLED || back == 2) { back = 3; break; }   // not the start of its run / the run reaches back to the previous sampled position
			r--; g--; back++;
		}
		if (back == 3) continue;
		const int lim = ml - r < sl - g ? ml - r : sl - g;
		int l = 0;
		while (l < lim)
		{
			u64 x = kb_rf_bits(w.mcode, r + l) ^ kb_rf_bits(w.wcode, g + l); x = (x | (x >> 1)) & M5;
			const int same = x ? (int)KB_CLZLL(x) >> 1 : 32;
			l += same;
			if (same < 32) break;
		}
		if (l > lim) l = lim;
		if (l < 10) continue;
		const u32 slot = KB_ATOMIC_ADD(&w.npairs, 1u);
		if (slot < (u32)KB_RF_PAIRS) { KbSeg sg; sg.simple = 1; sg.rpos = (i32)r; sg.gpos = (i64)g; sg.rlen = sg.glen = l; w.pairs[slot] = sg; }
	}
}
// all lanes: the left-maximal exact matches of >= 10 bases (kb_rj_pairs on packed words), in two phases.
// Scan: positions are dealt to the lanes in groups of 8 (one 32-base fetch serves the group's eight 8-mers out of a register); the
// 16 K-bit filter turns ~99 % of them away after one shared-memory load, the others are only noted in a short list. A lane that
// probed right away held the other 31 up for ~100 instructions whenever any lane passed (ncu r24: the scan ran at 16 of 32 lanes).
// Probe: the noted positions are dealt round-robin, so the ~150 consecutive positions of a real copy of the mate spread over all
// lanes. Positions beyond the list's capacity are probed on the spot.
// Sampling (r32): a run of >= 10 bases holds the 8-mers at >= 3 consecutive window positions, so every third position (g = 0, 3, 6 ...)
// meets each such run at least once, at most two bases behind its start: a third of the filter tests, the same set of runs
// (kb_rf_probe<true> walks back to the start and leaves the run to the earlier sampled position when there is one). One 32-base fetch
// serves nine sampled positions. KB_RF_STRIDE=1 scans every position (the r24 schedule) for the A/B.
KB_HD void kb_rf_scan(KbRescueFast& w, int lane)
{
	if (!w.ok || w.dirty) return;
	const int npos = w.slen - 7;
	if (w.stride == 3)
	{
		const int ngroups = (npos + 26) / 27;
		for (int grp = lane; grp < ngroups; grp += 32)
		{
			u64 bits = kb_rf_bits(w.wcode, grp * 27);
			const int gend = grp * 27 + 27 < npos ? grp * 27 + 27 : npos;
			for (int g = grp * 27; g < gend; g += 3, bits <<= 6)
			{
				const u32 id = (u32)(bits >> 48);
				const u32 h = kb_rf_fold(id);
				if (((w.filt[h >> 5] >> (h & 31)) & 1u) == 0u) continue;
				const u32 at = KB_ATOMIC_ADD(&w.ncand, 1u);
				if (at < w.cand_cap) w.cand[at] = (unsigned short)g; else kb_rf_probe<true>(w, g, id);
			}
		}
		return;
	}
	const int ngroups = (npos + 7) >> 3;
	for (int grp = lane; grp < ngroups; grp += 32)
	{
		u64 bits = kb_rf_bits(w.wcode, grp << 3);
		const int gend = (grp << 3) + 8 < npos ? (grp << 3) + 8 : npos;
		for (int g = grp << 3; g < gend; g++, bits <<= 2)
		{
			const u32 id = (u32)(bits >> 48);
			const u32 h = kb_rf_fold(id);
			if (((w.filt[h >> 5] >> (h & 31)) & 1u) == 0u) continue;
			const u32 at = KB_ATOMIC_ADD(&w.ncand, 1u);
			if (at < w.cand_cap) w.cand[at] = (unsigned short)g; else kb_rf_probe<false>(w, g, id);
		}
	}
}
KB_HD void kb_rf_pairs(KbRescueFast& w, int lane)
{
	if (!w.ok || w.dirty) return;
	const int n = (int)(w.ncand < w.cand_cap ? w.ncand : w.cand_cap);
	for (int i = lane; i < n; i += 32)
	{
		const int g = (int)w.cand[i]; const u32 id = (u32)(kb_rf_bits(w.wcode, g) >> 48);
		if (w.stride == 3) kb_rf_probe<true>(w, g, id); else kb_rf_probe<false>(w, g, id);
	}
}
// lane 0: the task's result (kb_rt_end), or the task's place on the slow list
KB_HD void kb_rf_end(const KbParams& pm, const KbBatchDev& bt, KbRescueFast& w)
{
	if (!w.ok || w.dirty || w.npairs > (u32)KB_RF_PAIRS) { const u32 slot = KB_ATOMIC_ADD(&bt.counters[29], 1u); bt.rslow[slot] = w.task; return; }
	const int np = (int)w.npairs;
	kb_sort_segs<false>(w.pairs, np);
	KbCand c;
	const int score = kb_rescue_cluster(pm, bt, w.left, w.pairs, np, &c);
	KbRTask* t = &bt.rtasks[w.task];
	t->score = score; t->diff = c.diff; t->seg_start = c.seg_start; t->nseg = c.nseg;
}

// ---- task-parallel rescue ----------------------------------------------------------------------------
// The anchors a rescue job visits, their reference windows and the thresholds all follow from the candidates the pair had
// BEFORE rescue (scores sc1/sc2, thr, EstDistance): what one window yields never depends on what another window appended.
// Only the bookkeeping is sequential (a rescued candidate takes the next free slot of its read). So the job is cut in three:
//   kb_rescue_plan   (thread per job)    strategy, then one task per (side, anchor) that has a usable window, in the reference's order
//   k_rescue_win     (block per task)    8-mer index of the mate, exact runs against the window, best diagonal cluster -> task result
//   kb_rescue_commit (thread per job)    walks the job's tasks in order and appends / cross-links exactly as :125-130,:160-165 do
// A pair with 100 anchors in a repeat family is then 100 blocks' work instead of one block's 100 windows in a row (the old
// kernel lasted as long as its longest job: 2.6 ms of an 11 ms step on the 100 Mbp repeat-rich index, ncu r11syn/r14 A/B).
KB_HD int kb_rescue_strategy(int sc1, int sc2, int l1, int l2)   // AlignmentRescue.cpp:86-93
{
	if (sc1 == 0 && sc2 == 0) return 0;
	if (sc1 < (int)(l1 * 0.1) && sc2 < (int)(l2 * 0.1)) return 4;
	if (sc1 > sc2 && sc1 - sc2 > 50) return 1;
	if (sc2 > sc1 && sc2 - sc1 > 50) return 2;
	return 3;
}
// the reference window of one anchor (:99-112 / :134-147); false when the anchor is skipped
KB_HD bool kb_rescue_window(const KbIndexDev& ix, int side, const KbCand& c, int est, int l2, int ml, i64* left_out, int* slen_out)
{
	i64 left, right; int cid, e;
	if (side == 0)
	{
		left = c.diff; right = c.diff + est + l2;
		e = kb_chr_lookup(ix, left); if (e >= ix.n_ends) return false;
		cid = ix.end_chr[e];
		if (right < ix.G && right > ix.chr_fwd[cid]) right = ix.chr_fwd[cid] - 1;
		else if (right >= ix.G && right > ix.chr_rev[cid]) right = ix.chr_rev[cid] - 1;
	}
	else
	{
		left = c.diff - est; right = c.diff + l2;
		e = kb_chr_lookup(ix, right); if (e >= ix.n_ends) return false;
		cid = ix.end_chr[e];
		if (left < ix.G && left < ix.chr_fwd[cid] - ix.chr_len[cid]) left = ix.chr_fwd[cid] - ix.chr_len[cid] + 1;
		else if (right >= ix.G && left < ix.chr_rev[cid] - ix.chr_len[cid]) left = ix.chr_rev[cid] - ix.chr_len[cid] + 1;
	}
	const int slen = (int)(right - left);
	if (slen < ml) return false;
	*left_out = left; *slen_out = slen;
	return true;
}
// thread per rescue job: the job's tasks, contiguous and in visiting order
KB_HD void kb_rescue_plan(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int k)
{
	const int p = bt.rescue_list[k], ra = 2 * p, rb = ra + 1;
	const int n1o = bt.n_cands[ra], n2o = bt.n_cands[rb];
	const int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
	const KbCand* a = bt.cands + bt.cand_off[ra]; const KbCand* b = bt.cands + bt.cand_off[rb];
	const int sc1 = kb_top_score(a, n1o), sc2 = kb_top_score(b, n2o);
	const int strategy = kb_rescue_strategy(sc1, sc2, l1, l2);
	bt.rjob_first[k] = 0; bt.rjob_count[k] = 0;
	if (strategy == 0 || strategy == 4) return;
	int est = bt.est[p]; if (est > pm.max_insert) est = pm.max_insert;
	u32 first = 0; int count = 0;
	for (int pass = 0; pass < 2; pass++)   // count, then fill
	{
		int w = 0;
		for (int side = 0; side < 2; side++)
		{
			if (side == 0 && !(strategy == 1 || strategy == 3)) continue;
			if (side == 1 && !(strategy == 2 || strategy == 3)) continue;
			const int top = side == 0 ? sc1 : sc2, thr = top - 30 < 50 ? 50 : top - 30;
			const int n = side == 0 ? n1o : n2o, ml = side == 0 ? l2 : l1;
			const KbCand* v = side == 0 ? a : b;
			for (int i = 0; i < n; i++)
			{
				if (v[i].score < thr) continue;
				i64 left; int slen;
				if (!kb_rescue_window(ix, side, v[i], est, l2, ml, &left, &slen)) continue;
				if (pass == 1 && (u64)first + (u64)w < (u64)bt.cap_rtasks)
				{
					KbRTask t; t.job = (u32)k; t.side = side; t.idx = i; t.slen = slen; t.left = left; t.score = 0; t.nseg = 0; t.diff = 0; t.seg_start = 0; t.pad = 0;
					bt.rtasks[first + (u32)w] = t;
				}
				w++;
			}
		}
		if (pass == 0)
		{
			count = w; if (count == 0) break;
			first = KB_ALLOC(&bt.counters[27], (u32)count);
			if ((u64)first + (u64)count > (u64)bt.cap_rtasks) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_RESCUE); return; }
		}
	}
	bt.rjob_first[k] = first; bt.rjob_count[k] = (u32)count;
}
// thread 0 of a task's block: the job record for the window phases (kb_rj_window .. kb_rj_pairs are shared with the description above)
KB_HD void kb_rt_begin(const KbIndexDev& ix, const KbBatchDev& bt, KbRescueJob* j, KbArena& ar, const KbRTask& t)
{
	j->p = bt.rescue_list[t.job]; j->ra = 2 * j->p; j->rb = j->ra + 1;
	j->l1 = (int)(bt.seq_off[j->ra + 1] - bt.seq_off[j->ra]); j->l2 = (int)(bt.seq_off[j->rb + 1] - bt.seq_off[j->rb]);
	j->side = t.side; j->idx = t.idx; j->done = 0; j->ovf = 0; j->npairs = 0; j->reindex = 1;
	if (t.side == 0) { j->mate = bt.seq + bt.seq_off[j->rb]; j->ml = j->l2; } else { j->mate = bt.seq + bt.seq_off[j->ra]; j->ml = j->l1; }
	const int lm = j->ml;
	ar.used = 0; ar.ovf = false;
	j->wm = (u32*)ar.alloc((u64)lm * 4);
	int hs = 256; while (hs < 2 * lm) hs <<= 1;
	j->hmask = hs - 1;
	j->hkey = (u32*)ar.alloc((u64)hs * 4); j->hhead = (i32*)ar.alloc((u64)hs * 4); j->hnext = (i32*)ar.alloc((u64)lm * 4);
	const int slen = t.slen;
	j->left = t.left; j->slen = slen; j->clean = (t.left >= 0 && t.left + slen <= ix.G2) ? 1 : 0;
	j->win = (u8*)ar.alloc((u64)slen); j->ww = (u32*)ar.alloc((u64)slen * 4);
	u64 worst = (u64)((j->ml < slen ? j->ml : slen) / 9 + 2) * (u64)(j->ml + slen);
	u64 room = ar.cap > ar.used ? (ar.cap - ar.used) / (2 * sizeof(KbSeg)) : 0;
	j->cap_pairs = (int)(worst < room ? worst : room);
	j->pairs = (KbSeg*)ar.alloc((u64)(j->cap_pairs > 0 ? j->cap_pairs : 1) * sizeof(KbSeg));
	if (ar.ovf) j->ovf = 1;
}
// thread 0: order the runs like GenerateSimplePairsFromCommonKmers does, pick the best diagonal cluster, record it in the task
KB_HD void kb_rt_end(const KbParams& pm, const KbBatchDev& bt, KbRescueJob* j, KbRTask* t)
{
	if (j->ovf) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH); return; }
	int np = (int)j->npairs;
	kb_sort_segs<false>(j->pairs, np);
	KbCand c;
	const int score = kb_rescue_cluster(pm, bt, j->left, j->pairs, np, &c);
	t->score = score; t->diff = c.diff; t->seg_start = c.seg_start; t->nseg = c.nseg;
}
// thread per rescue job: AlignmentRescue.cpp:125-130 / :160-165 over the job's tasks, then what follows in ReadMapping (Mapping.cpp:561-563)
KB_HD void kb_rescue_commit(const KbParams& pm, const KbBatchDev& bt, int k, u32* n_attempted)
{
	const int p = bt.rescue_list[k], ra = 2 * p, rb = ra + 1;
	const int n1o = bt.n_cands[ra], n2o = bt.n_cands[rb];
	const int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
	KbCand* a = bt.cands + bt.cand_off[ra]; KbCand* b = bt.cands + bt.cand_off[rb];
	const int sc1 = kb_top_score(a, n1o), sc2 = kb_top_score(b, n2o);
	const int strategy = kb_rescue_strategy(sc1, sc2, l1, l2);
	const bool attempted = !(strategy == 0 || strategy == 4);
	int n1 = n1o, n2 = n2o; bool mated = false;
	const u32 first = bt.rjob_first[k], count = bt.rjob_count[k];
	for (u32 q = 0; q < count; q++)
	{
		const KbRTask t = bt.rtasks[first + q];
		KbCand c; c.score = t.score; c.diff = t.diff; c.seg_start = t.seg_start; c.nseg = t.nseg; c.mate = t.idx;
		if (t.side == 0) { if (t.score > sc2 && n2 < bt.cand_cap[rb]) { mated = true; a[t.idx].mate = n2; b[n2++] = c; } }
		else if (t.score > sc1 && n1 < bt.cand_cap[ra]) { mated = true; b[t.idx].mate = n1; a[n1++] = c; }
	}
	bt.n_cands[ra] = n1; bt.n_cands[rb] = n2;
	if (attempted)
	{
		*n_attempted += 1;   // counters[7], added up per warp by the kernel (one same-address atomic per job was most of k_rescue_commit)
		KbPairStat& st = bt.pstat[p]; int est = bt.est[p];
		if (est >= pm.max_insert) { if (st.est_lo < pm.max_insert) st.est_lo = pm.max_insert; }
		else { st.est_lo = est; st.est_hi = est; }
	}
	if (mated) kb_keep_mated(a, n1, b, n2);
	kb_prune(pm, a, n1); kb_prune(pm, b, n2);
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

// r1 / r2 and p1 / p2 are the pair's results and reports where the caller holds them: in the arenas, or copies in the thread's local memory
KB_HD void kb_finalize_pair_on(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int p, KbReadRes& r1, KbReport* p1, KbReadRes& r2, KbReport* p2, int l1, int l2)
{
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
KB_HD void kb_finalize_pair(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int p)
{
	int ra = 2 * p, rb = ra + 1;
	KbReadRes& r1 = bt.res[ra]; KbReadRes& r2 = bt.res[rb];
	KbReport* p1 = bt.reports + r1.rep_off; KbReport* p2 = bt.reports + r2.rep_off;
	int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
	kb_finalize_pair_on(ix, pm, bt, p, r1, p1, r2, p2, l1, l2);
}

KB_HD void kb_finalize_single(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r)
{
	KbReadRes& rd = bt.res[r];
	kb_flag_single(rd, bt.reports + rd.rep_off);
	kb_mapq(ix, pm, rd, (int)(bt.seq_off[r + 1] - bt.seq_off[r]));
}

#endif
