// Per-thread bodies of the pipeline kernels (k_cand_pair, k_cand_pacbio, k_rescue, k_report, k_finalize in kb_api.cu).
// Kept as KB_HD functions so that tests/emul can run the identical logic on the host next to the oracle.
#ifndef KB_STAGES_CUH
#define KB_STAGES_CUH
#include "../../include/kart_b200.h"
#include "kb_pair.cuh"

// seeds -> sorted seeds -> candidates -> pairing (one thread per pair, or per read when not paired).
// A read inside a repeat family brings dozens to hundreds of seeds (up to 50 per search); sorting those in one thread while the
// other 31 pairs of the warp wait was 36 % of k_cand_pair's samples at 1.6 lanes (ncu r17, C3). Such items (more than
// KB_CAND_HEAVY seeds on either read) are therefore only set up here and put on the heavy list; k_cand_heavy gives each of them
// a warp that sorts cooperatively (kb_wsort_*) and finishes the rest on one lane without holding anybody up.
#define KB_CAND_HEAVY 24
// part 1: candidate slots and the pair's statistics record. False: nothing more to do (overflow, or pacbio's own kernel follows).
KB_HD bool kb_cand_setup(const KbParams& pm, const KbBatchDev& bt, int t)
{
	if (pm.paired)
	{
		int ra = 2 * t, rb = ra + 1;
		int s1 = bt.n_seeds[ra], s2 = bt.n_seeds[rb];
		int cap = s1 + s2 + 1;
		u32 off = KB_ALLOC(&bt.counters[1], (u32)(2 * cap));
		bt.cand_off[ra] = off; bt.cand_off[rb] = off + cap; bt.cand_cap[ra] = cap; bt.cand_cap[rb] = cap;
		bt.n_cands[ra] = 0; bt.n_cands[rb] = 0;
		KbPairStat st; st.counted = 0; st.absdist = 0; st.est_lo = -2147483647 - 1; st.est_hi = 2147483647; bt.pstat[t] = st;
		if ((u64)off + 2ull * cap > (u64)bt.cap_cands) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_CANDS); bt.cand_off[ra] = 0; bt.cand_off[rb] = 0; return false; }
		if (bt.counters[3] & (KB_OVF_SEEDS | KB_OVF_HITS)) return false;
		return true;
	}
	int s1 = bt.n_seeds[t], cap = s1 + 1;
	u32 off = KB_ALLOC(&bt.counters[1], (u32)cap);
	bt.cand_off[t] = off; bt.cand_cap[t] = cap; bt.n_cands[t] = 0;
	if ((u64)off + (u64)cap > (u64)bt.cap_cands) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_CANDS); bt.cand_off[t] = 0; return false; }
	if (bt.counters[3] & (KB_OVF_SEEDS | KB_OVF_HITS)) return false;
	if (pm.pacbio) return false;   // pacbio chaining needs scratch: done by k_cand_pacbio
	return true;
}
KB_HD bool kb_cand_is_heavy(const KbParams& pm, const KbBatchDev& bt, int t)
{
	if (pm.paired) return bt.n_seeds[2 * t] > KB_CAND_HEAVY || bt.n_seeds[2 * t + 1] > KB_CAND_HEAVY;
	return bt.n_seeds[t] > KB_CAND_HEAVY;
}
// part 3 (the seeds are (PosDiff,rPos)-sorted): candidates, pairing, pruning, rescue list
template <bool WIDE>
KB_HD void kb_cand_finish(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int t)
{
	if (pm.paired)
	{
		int ra = 2 * t, rb = ra + 1;
		int s1 = bt.n_seeds[ra], s2 = bt.n_seeds[rb], cap = bt.cand_cap[ra];
		KbSeg* v1 = bt.segs + bt.seed_off[ra]; KbSeg* v2 = bt.segs + bt.seed_off[rb];
		int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
		KbCand* a = bt.cands + bt.cand_off[ra]; KbCand* b = a + cap;
		int n1 = kb_cands_illumina<WIDE>(ix, pm, l1, v1, s1, bt.seed_off[ra], a, cap);
		int n2 = kb_cands_illumina<WIDE>(ix, pm, l2, v2, s2, bt.seed_off[rb], b, cap);
		bt.n_cands[ra] = n1; bt.n_cands[rb] = n2;
		i32 lo = bt.pstat[t].est_lo, hi = bt.pstat[t].est_hi;
		bool paired = kb_pair(pm, (i64)bt.est[t], a, n1, b, n2, &lo, &hi);
		bt.pstat[t].est_lo = lo; bt.pstat[t].est_hi = hi;
		if (paired) { kb_keep_mated(a, n1, b, n2); kb_prune(pm, a, n1); kb_prune(pm, b, n2); }
		else if (kb_top_score(a, n1) == 0 && kb_top_score(b, n2) == 0) { kb_prune(pm, a, n1); kb_prune(pm, b, n2); }   // AlignmentRescue.cpp:89
		else { u32 slot = KB_ALLOC(&bt.counters[4], 1u); bt.rescue_list[slot] = t; }
	}
	else
	{
		int s1 = bt.n_seeds[t];
		KbSeg* v = bt.segs + bt.seed_off[t];
		int l = (int)(bt.seq_off[t + 1] - bt.seq_off[t]);
		KbCand* a = bt.cands + bt.cand_off[t];
		int n = kb_cands_illumina<WIDE>(ix, pm, l, v, s1, bt.seed_off[t], a, bt.cand_cap[t]);
		bt.n_cands[t] = n;
		kb_prune(pm, a, n);
	}
}
KB_HD void kb_stage_cand_pair(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int t, bool split_heavy)
{
	if (t >= (pm.paired ? (bt.n_reads >> 1) : bt.n_reads)) return;
	if (!kb_cand_setup(pm, bt, t)) return;
	if (split_heavy && kb_cand_is_heavy(pm, bt, t)) { const u32 slot = KB_ALLOC(&bt.counters[15], 1u); bt.slow_list2[slot] = t; return; }
	if (pm.paired) { kb_sort_segs<false>(bt.segs + bt.seed_off[2 * t], bt.n_seeds[2 * t]); kb_sort_segs<false>(bt.segs + bt.seed_off[2 * t + 1], bt.n_seeds[2 * t + 1]); }
	else kb_sort_segs<false>(bt.segs + bt.seed_off[t], bt.n_seeds[t]);
	kb_cand_finish<false>(ix, pm, bt, t);
}

// ---- warp-cooperative (PosDiff,rPos) sort of one read's seeds (k_cand_heavy) ----
// Keys are packed into 64 bits (PosDiff biased into 43 bits, rPos in 20) and sorted with their index by a bitonic network in the
// warp's shared memory; the seeds are then permuted through `tmp`, a region of at least n entries nobody uses yet (the pair's
// candidate slots: a KbCand is as large as a KbSeg). Both orders are total on distinct seeds, so the result equals kb_sort_segs'.
#define KB_WSORT_MAX 1024
struct KbWarpSort { u64 key[KB_WSORT_MAX]; unsigned short idx[KB_WSORT_MAX]; KbSeg* v; KbSeg* tmp; int n, p, by_gpos; };
KB_HD void kb_wsort_begin(KbWarpSort& w, KbSeg* v, int n, KbSeg* tmp, bool by_gpos = false) { w.v = v; w.tmp = tmp; w.n = n; w.by_gpos = by_gpos ? 1 : 0; int p = 32; while (p < n) p <<= 1; w.p = p; }
// by_gpos: the (gPos,rPos) order of the pacbio path (AlignmentCandidates.cpp:17-21): gPos in the upper 44 bits, rPos (< 2^20) below
KB_HD void kb_wsort_load(KbWarpSort& w, int lane)
{
	for (int i = lane; i < w.p; i += 32)
	{
		u64 k = ~0ull;
		if (i < w.n)
		{
			const KbSeg s = w.v[i];
			k = w.by_gpos ? (((u64)s.gpos << 20) | (u64)(u32)s.rpos) : (((u64)((s.gpos - (i64)s.rpos) + ((i64)1 << 42)) << 20) | (u64)(u32)s.rpos);
		}
		w.key[i] = k; w.idx[i] = (unsigned short)i;
	}
}
KB_HD void kb_wsort_step(KbWarpSort& w, int k, int j, int lane)
{
	for (int i = lane; i < w.p; i += 32)
	{
		const int o = i ^ j;
		if (o <= i) continue;
		const u64 a = w.key[i], b = w.key[o];
		if ((a > b) == ((i & k) == 0)) { w.key[i] = b; w.key[o] = a; const unsigned short t = w.idx[i]; w.idx[i] = w.idx[o]; w.idx[o] = t; }
	}
}
KB_HD void kb_wsort_gather(KbWarpSort& w, int lane) { for (int i = lane; i < w.n; i += 32) w.tmp[i] = w.v[w.idx[i]]; }
KB_HD void kb_wsort_scatter(KbWarpSort& w, int lane) { for (int i = lane; i < w.n; i += 32) w.v[i] = w.tmp[i]; }
// which seed lists of heavy item t the warp sorts: list 0 / 1, the list and its length; false when there is no such list
KB_HD bool kb_cand_heavy_list(const KbParams& pm, const KbBatchDev& bt, int t, int which, KbSeg** v, int* n, KbSeg** tmp)
{
	if (which > (pm.paired ? 1 : 0)) return false;
	const int r = pm.paired ? 2 * t + which : t;
	*v = bt.segs + bt.seed_off[r]; *n = bt.n_seeds[r];
	*tmp = reinterpret_cast<KbSeg*>(bt.cands + bt.cand_off[pm.paired ? 2 * t : t]);   // 2 x (s1 + s2 + 1) (or s1 + 1) unused candidate slots
	return true;
}

// Private arena of thread `tid` of an arena kernel launched with `nth` threads. The scratch buffer is sized for the worst
// case per thread (a 3000 x 3000 traceback); the kernels that run one thread per read need far less than that per thread,
// so they cut the same buffer into slices of `need` bytes (a bound computed on the device from the batch's own maxima) and
// get that many more threads. Returns the number of threads that own a slice; the others must idle.
KB_HD int kb_thread_arena(const KbBatchDev& bt, int tid, int nth, u64 need, KbArena* ar)
{
	need = (need + 255) & ~(u64)255;
	const u64 total = bt.scratch_per_thread * (u64)bt.scratch_threads;
	u64 fit = total / need; if (fit > (u64)nth) fit = (u64)nth;
	ar->base = bt.scratch + (u64)tid * need; ar->used = 0; ar->cap = need; ar->ovf = false;
	return (int)fit;
}

// sorted: k_cand_pacbio_sort has put every read's seeds into (gPos,rPos) order already
KB_HD void kb_stage_cand_pacbio(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int tid, int nth, bool sorted)
{
	if (bt.counters[3]) return;
	KbArena ar; nth = kb_thread_arena(bt, tid, nth, ((u64)bt.counters[5] + 2) * 32ull + 256ull, &ar);   // taken[] + a copy of the read's seeds
	if (tid >= nth) return;
	for (int r = tid; r < bt.n_reads; r += nth)
	{
		int n = bt.n_seeds[r];
		KbSeg* v = bt.segs + bt.seed_off[r];
		if (!sorted) kb_sort_segs<true>(v, n);
		ar.used = 0;
		u8* taken = (u8*)ar.alloc((u64)n + 1); KbSeg* tmp = (KbSeg*)ar.alloc((u64)(n + 1) * sizeof(KbSeg));
		if (ar.ovf) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH); return; }
		KbCand* a = bt.cands + bt.cand_off[r];
		int nc = kb_cands_pacbio(bt, v, n, taken, tmp, a, bt.cand_cap[r]);
		bt.n_cands[r] = nc;
		kb_prune(pm, a, nc);
	}
}

// phase A of the report stage: segments of every surviving candidate. One thread per read, local memory only; reads with a
// candidate of more than KB_SEG_FAST seeds go to the slow list and are done by the arena version (grid-stride, private arena)
KB_HD void kb_stage_segments(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r)
{
	if (r >= bt.n_reads) return;
	if (bt.counters[3]) return;
	if (!kb_segments_read(ix, pm, bt, r, nullptr)) { u32 slot = KB_ALLOC(&bt.counters[12], 1u); bt.slow_list[slot] = r; }
}
KB_HD void kb_stage_segments_slow(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int tid, int nth)
{
	if (bt.counters[3]) return;
	KbArena ar; nth = kb_thread_arena(bt, tid, nth, (u64)(bt.counters[5] > 64u ? bt.counters[5] : 64u) * 128ull + 1024ull, &ar);   // seeds + 2 x segments + order of one candidate
	if (tid >= nth) return;
	const int count = (int)bt.counters[12];
	for (int k = tid; k < count; k += nth)
	{
		ar.used = 0;
		kb_segments_read(ix, pm, bt, bt.slow_list[k], &ar);
		if (ar.ovf) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH); return; }
	}
}

// phase C: reports; same split (a cigar of more than KB_CIG_FAST elements sends the read to the arena version)
KB_HD void kb_stage_assemble(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int r)
{
	if (r >= bt.n_reads) return;
	if (bt.counters[3]) return;
	if (!kb_assemble_read(ix, pm, bt, r, nullptr)) { u32 slot = KB_ALLOC(&bt.counters[13], 1u); bt.slow_list2[slot] = r; }
}
KB_HD void kb_stage_assemble_slow(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, int tid, int nth)
{
	if (bt.counters[3]) return;
	KbArena ar; nth = kb_thread_arena(bt, tid, nth, 16ull * (u64)bt.max_rlen + 64ull * (u64)(bt.counters[5] > 64u ? bt.counters[5] : 64u) + 2048ull, &ar);   // cigar elements of one candidate (kb_assemble_read)
	if (tid >= nth) return;
	const int count = (int)bt.counters[13];
	for (int k = tid; k < count; k += nth)
	{
		ar.used = 0;
		kb_assemble_read(ix, pm, bt, bt.slow_list2[k], &ar);
		if (ar.ovf) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_SCRATCH); return; }
	}
}

#define KB_FIN_LOCAL 4   // reports per read that k_finalize holds in local memory (see kb_stage_finalize)
// what one SAM line needs (OutputPairedAlignments / OutputSingledAlignments, src/Mapping.cpp:177-315)
KB_HD void kb_fill_aln(kb_aln_t& o, const KbReadRes& rd, const KbReport* rep, const KbReport* mate_rep, bool mate_ok, int tlen)
{
	o.score = rd.score; o.sub_score = rd.sub; o.mapq = rd.mapq; o.mate_pos = -1; o.tlen = 0; o.pos = 0; o.chr = 0; o.cig_off = 0; o.cig_len = 0; o.fwd = 1;
	if (rd.score == 0) { o.kind = 0; o.flag = rep[0].flag; return; }
	const KbReport& a = rep[rd.best];
	if (a.aln <= 0) { o.kind = 2; o.flag = 0; return; }
	o.kind = 1; o.flag = a.flag; o.chr = a.chr; o.pos = a.pos; o.cig_off = a.cig_off; o.cig_len = a.cig_len; o.fwd = a.fwd;
	if (mate_ok) { o.mate_pos = mate_rep->pos; o.tlen = tlen; }
}

// -m: the further lines of a read = the reports after iBestAlnCanIdx, in AlnReportArr order, that the loops of
// OutputPairedAlignments (AlnScore > 0, src/Mapping.cpp:194-223,242-263) / OutputSingledAlignments (AlnScore == score, :289-308) print.
// which: 0 single-end, 1 first mate, 2 second mate. l1/l2: lengths of the two mates.
KB_HD void kb_emit_extra(const KbBatchDev& bt, int r, const KbReadRes& rd, const KbReport* rep, const KbReport* mrep, int l1, int l2, int which)
{
	if (rd.score == 0) return;
	int cnt = 0;
	for (int i = rd.best + 1; i < rd.ncan; i++) cnt += (which ? rep[i].aln > 0 : rep[i].aln == rd.score) ? 1 : 0;
	if (cnt == 0) return;
	u32 at = KB_ALLOC(&bt.counters[14], (u32)cnt);
	if ((u64)at + (u64)cnt > (u64)bt.cap_extra) { KB_ATOMIC_OR(&bt.counters[3], (u32)KB_OVF_EXTRA); return; }
	u32 k = 0;
	for (int i = rd.best + 1; i < rd.ncan; i++)
	{
		const KbReport& a = rep[i];
		if (!(which ? a.aln > 0 : a.aln == rd.score)) continue;
		kb_extra_t& e = bt.extra[at + k]; e.read = (u32)r; e.rank = k; k++;
		kb_aln_t& o = e.aln;
		o.score = rd.score; o.sub_score = rd.sub; o.mapq = rd.mapq; o.mate_pos = -1; o.tlen = 0;
		o.kind = 1; o.flag = a.flag; o.chr = a.chr; o.pos = a.pos; o.cig_off = a.cig_off; o.cig_len = a.cig_len; o.fwd = a.fwd;
		int j = which ? a.mate : -1;
		if (j != -1 && mrep[j].aln > 0)
		{
			o.mate_pos = mrep[j].pos;
			o.tlen = which == 1 ? (int)(mrep[j].pos - a.pos + (a.fwd ? l2 : 0 - l1)) : 0 - (int)(a.pos - mrep[j].pos + (mrep[j].fwd ? l2 : 0 - l1));
		}
	}
}

// aln: where this item's records go -- aln[0] and aln[1] for a pair, aln[0] for a single read (the kernel stages them in shared memory)
KB_HD void kb_stage_finalize(const KbIndexDev& ix, const KbParams& pm, const KbBatchDev& bt, kb_aln_t* aln, int t)
{
	if (bt.counters[3]) return;
	if (pm.paired)
	{
		if (t >= (bt.n_reads >> 1)) return;
		int ra = 2 * t, rb = ra + 1;
		int l1 = (int)(bt.seq_off[ra + 1] - bt.seq_off[ra]), l2 = (int)(bt.seq_off[rb + 1] - bt.seq_off[rb]);
		// The pair's reports are 40-byte records scattered by candidate; settling, flagging and filling walk them back and forth, each step
		// a dependent load behind the previous store (ncu r29: 326 stall cycles per issued instruction, 16 M instructions per million reads).
		// Pairs with at most KB_FIN_LOCAL reports per read (nearly all) are therefore fetched ONCE, with independent loads, into the thread's
		// local memory, finished there with the same functions, and written back.
		KbReadRes c1 = bt.res[ra], c2 = bt.res[rb];
		if (bt.fin_local && !pm.multihit && c1.ncan <= KB_FIN_LOCAL && c2.ncan <= KB_FIN_LOCAL)
		{
			KbReport q1[KB_FIN_LOCAL], q2[KB_FIN_LOCAL];
			KbReport* g1 = bt.reports + c1.rep_off; KbReport* g2 = bt.reports + c2.rep_off;
			const int n1 = c1.ncan > 1 ? c1.ncan : 1, n2 = c2.ncan > 1 ? c2.ncan : 1;   // an unmapped read still owns the report that carries its flag
			for (int i = 0; i < KB_FIN_LOCAL; i++) { if (i < n1) q1[i] = g1[i]; if (i < n2) q2[i] = g2[i]; }
			kb_finalize_pair_on(ix, pm, bt, t, c1, q1, c2, q2, l1, l2);
			{
				const KbReport& a = q1[c1.best]; int j = a.mate; bool ok = c1.score > 0 && a.aln > 0 && j != -1 && q2[j].aln > 0;
				int dist = ok ? (int)(q2[j].pos - a.pos + (a.fwd ? l2 : 0 - l1)) : 0;
				kb_fill_aln(aln[0], c1, q1, ok ? &q2[j] : nullptr, ok, dist);
			}
			{
				const KbReport& b = q2[c2.best]; int i = b.mate; bool ok = c2.score > 0 && b.aln > 0 && i != -1 && q1[i].aln > 0;
				int dist = ok ? 0 - (int)(b.pos - q1[i].pos + (q1[i].fwd ? l2 : 0 - l1)) : 0;
				kb_fill_aln(aln[1], c2, q2, ok ? &q1[i] : nullptr, ok, dist);
			}
			bt.res[ra] = c1; bt.res[rb] = c2;
			for (int i = 0; i < KB_FIN_LOCAL; i++) { if (i < n1) g1[i] = q1[i]; if (i < n2) g2[i] = q2[i]; }
			return;
		}
		kb_finalize_pair(ix, pm, bt, t);
		const KbReadRes& r1 = bt.res[ra]; const KbReadRes& r2 = bt.res[rb];
		const KbReport* p1 = bt.reports + r1.rep_off; const KbReport* p2 = bt.reports + r2.rep_off;
		{   // read 1 line (:194-221)
			const KbReport& a = p1[r1.best]; int j = a.mate; bool ok = r1.score > 0 && a.aln > 0 && j != -1 && p2[j].aln > 0;
			int dist = ok ? (int)(p2[j].pos - a.pos + (a.fwd ? l2 : 0 - l1)) : 0;
			kb_fill_aln(aln[0], r1, p1, ok ? &p2[j] : nullptr, ok, dist);
		}
		{   // read 2 line (:242-261)
			const KbReport& b = p2[r2.best]; int i = b.mate; bool ok = r2.score > 0 && b.aln > 0 && i != -1 && p1[i].aln > 0;
			int dist = ok ? 0 - (int)(b.pos - p1[i].pos + (p1[i].fwd ? l2 : 0 - l1)) : 0;
			kb_fill_aln(aln[1], r2, p2, ok ? &p1[i] : nullptr, ok, dist);
		}
		if (pm.multihit) { kb_emit_extra(bt, ra, r1, p1, p2, l1, l2, 1); kb_emit_extra(bt, rb, r2, p2, p1, l1, l2, 2); }
	}
	else
	{
		if (t >= bt.n_reads) return;
		kb_finalize_single(ix, pm, bt, t);
		const KbReadRes& rd = bt.res[t]; const KbReport* rep = bt.reports + rd.rep_off;
		kb_fill_aln(aln[0], rd, rep, nullptr, false, 0);
		if (rd.score > 0 && rep[rd.best].aln != rd.score) aln[0].kind = 2;   // OutputSingledAlignments prints reports with AlnScore == score (:293)
		if (pm.multihit) kb_emit_extra(bt, t, rd, rep, nullptr, 0, 0, 0);
	}
}


#endif
