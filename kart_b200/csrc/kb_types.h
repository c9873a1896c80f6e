// Shared POD types for the kart_b200 CUDA path (host + device).
//
// Everything the kernels touch lives in HBM in these layouts:
//   * FM-index: re-blocked at upload from the BWA .bwt layout (reference src/BWT_Index/bwtindex.c:53-75:
//     64-byte blocks = 4 x u64 counts + 128 symbols) into 32-byte blocks = ONE DRAM sector per Occ query:
//     [u32 cntA,cntC,cntG,cntT][4 x u32 = 64 symbols, 16 per word, MSB first]. Results are identical by construction.
//   * sampled SA exactly as in .sa (every 32nd row, sa[0] = ~0), optionally expanded to a full SA.
//   * reference: the forward strand 2-bit .pac bytes; the reverse-complement half of Kart's 2G text
//     (src/bwt_index.cpp:194-213) is computed on the fly.
//   * per-batch worklists: hits -> seeds -> candidates -> reports -> cigar arena (bump-allocated, capacity checked).
#ifndef KB_TYPES_H
#define KB_TYPES_H
#include <stdint.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#define KB_D __device__ __forceinline__
#else
#define KB_HD inline
#define KB_D inline
#define KB_NCOUNTERS 32
#define KB_NW_CLASSES 7  // nw_alignment size classes: longer side <= 8, 16, 24, 32 (one register tile), <= 64, <= 128 (column tiles of 32), else warp wavefront
#define KB_NW_TMAX 128   // largest side a single thread solves

#endif

struct KbParams
{
	i32 min_seed;      // MinSeedLength, src/Mapping.cpp:645
	i32 max_gaps;      // -g, src/main.cpp:92
	i32 max_insert;    // MaxInsertSize 1500, src/main.cpp:96
	i32 pacbio;        // -pacbio
	i32 multihit;      // -m
	i32 paired;        // reads come as (mate1, revcomp(mate2)) pairs
};

// seeding table: state of BWT_Search after the first K bases (x2 > 0), or the length at which the search died (x2 == 0).
// 16 bytes per entry, two entries per DRAM sector: a = x0 | (x2 & 0xFFFFFF) << 40, b = x1 | (x2 >> 24) << 40 (row numbers below 2^40);
// a dead search has a = 0 and the length in b. KbKtabE is the unpacked form.
struct __attribute__((aligned(16))) KbKtab { u64 a, b; };
struct KbKtabE { u64 x0, x1, x2; u32 flen; };

struct KbIndexDev
{
	u64 primary, L2[5], seq_len;
	const uint32_t* occ;     // 8 words per 64-row block
	u64 n_blocks;
	const u64* sa;           // sampled SA (.sa layout), sa[0] = ~0
	u64 n_sa;
	i32 sa_intv;
	const u64* sa_full;      // optional full SA (NULL when not expanded)
	const KbKtab* ktab;      // 4^ktab_k entries, indexed by the 2-bit codes of the first ktab_k bases (first base most significant)
	i32 ktab_k;
	const u8* pac;           // forward strand, 2 bit / base, MSB first
	const u64* ref64;        // the same bits as big-endian 64-bit words: word w = bases 32w..32w+31, base i at bits 62-2(i&31)
	i64 G, G2;               // GenomeSize, TwoGenomeSize
	i32 n_chr, n_ends;
	const i64* end_key;      // ChrLocMap keys (last coordinate of each chromosome on each strand), ascending
	const i32* end_chr;      // ChrLocMap values
	const uint16_t* end_tab; // end_tab[b] = index of the first key >= b << end_shift (NULL: plain binary search); narrows kb_chr_lookup to a bucket
	i32 end_shift;
	const i64* chr_fwd;      // Chromosome_t::FowardLocation
	const i64* chr_rev;      // Chromosome_t::ReverseLocation
	const i64* chr_len;
	const u8* mapq_lut;      // [score][diff-1], diff = 1..5 ; built on the host with the reference expression
	i32 mapq_lut_scores;
	i32 ld_hint;             // seeding kernels: bit 0 = Occ blocks, table and SA entries bypass L1 allocation (kb_load_blk); bit 1 = kb_unique_tail without its fast path
};

// 32 read characters: nt4 codes (2 bit, MSB first, 0 where the character is no base), n4 = "nt4 code is 4" and
// bad = "not one of the upper-case letters ACGT" (bit 31-i for character i; set beyond the end of the read)
struct __attribute__((aligned(16))) KbPk { u64 code; u32 n4, bad; };
struct KbHit { u64 x0; u32 rpos; u32 len_freq; };                 // len << 8 | freq   (freq <= 50)
struct KbSeg { i64 gpos; i32 rpos; i32 rlen; i32 glen; i32 simple; };
struct KbCand { i64 diff; i32 score; i32 mate; u32 seg_start; i32 nseg; };
struct KbReport { i64 pos; i32 aln; i32 flag; i32 mate; i32 chr; u32 cig_off; i32 cig_len; i32 fwd; i32 pad; };
struct KbReadRes { i32 score, sub, mapq, ncan, best; u32 rep_off; };
struct KbPairStat { i32 counted, absdist, est_lo, est_hi; };
// one segment of a candidate after IdentifyNormalPairs, with how it is resolved (info) and a class-specific value (aux)
struct KbSegX { KbSeg s; u32 info; u32 aux; };
enum { KB_SEG_SKIP = 0, KB_SEG_SIMPLE = 1, KB_SEG_QUICK = 2 /* aux = score */, KB_SEG_GAP = 3, KB_SEG_SOFT = 4, KB_SEG_JOB = 5 /* aux = job id */, KB_SEG_ONE = 6 /* 1x1, aux = identical? */ };
// one fragment pair that needs GenerateNormalPairAlignment (k-mer partition + NW); result = run list in the run arena
struct KbJob { i64 gpos; u32 read; i32 rpos, rlen, glen; u32 run_off; i32 nruns, ident, aligned; };
// one nw_alignment call: the sub-rectangle (r0,rl) x (g0,gl) of job `job`; its runs go to runs[out_off .. out_off + rl + gl), written from
// the end of that slice backwards. whole != 0: the piece is the entire job (no 8-mer partition): the solver finishes the job itself.
struct KbPiece { u32 job; i32 r0, rl, g0, gl; u32 out_off; u32 whole; u32 pad; };

// one reference window of a rescue job (kb_pair.cuh "task-parallel rescue"): the task, and what its block found there
struct KbRTask { u32 job; i32 side, idx, slen; i64 left; i32 score, nseg; i64 diff; u32 seg_start, pad; };

// status bits written by kernels (any non-zero value fails the batch loudly; capacities are then grown and the batch rerun)
enum { KB_OVF_SEEDS = 1, KB_OVF_CANDS = 2, KB_OVF_CIGAR = 4, KB_OVF_SCRATCH = 8, KB_OVF_HITS = 16, KB_OVF_RESCUE = 32, KB_OVF_NW = 64, KB_OVF_SEGX = 128, KB_OVF_JOBS = 256, KB_OVF_RUNS = 512, KB_OVF_EXTRA = 1024 };

// cigar op codes in the arena: len << 4 | op  (BAM numbering)
enum { KB_OP_M = 0, KB_OP_I = 1, KB_OP_D = 2, KB_OP_S = 4 };

// Everything one batch needs on the device. Arrays are device pointers.
struct KbBatchDev
{
	i32 n_reads;
	const u8* seq;           // concatenated read characters (mate 2 already reverse-complemented, src/GetData.cpp:125-135)
	const u64* seq_off;      // n_reads + 1
	const i32* est;          // per pair EstDistance (n_reads / 2 entries) when paired
	KbPk* pk; i32 pk_wpr;    // packed reads: read r starts at word (seq_off[r] >> 5) + r ; pk_wpr = words of the longest read
	// stage 1
	KbHit* hits; i32 max_hits;          // [n_reads][max_hits]
	i32* n_hits;                        // hits stored per read
	i32* n_seeds; u32* seed_off;        // seeds per read and their offset in `segs`
	KbSeg* segs; u32 cap_segs;
	// stage 2
	KbCand* cands; u32 cap_cands; i32* n_cands; u32* cand_off; i32* cand_cap;
	i32* rescue_list;                   // pair ids that need rescue
	u32* segx_slab;                     // {next, end} of the calling warp's reserved range of segx (shared memory; NULL: none). Set by k_segments
	KbRTask* rtasks; u32 cap_rtasks;    // rescue windows (cursor: counters[27]); a job's tasks are rtasks[rjob_first[k] .. + rjob_count[k])
	u32* rjob_first; u32* rjob_count;
	u32* rslow;                         // rescue windows the warp-per-window fast path handed back (count: counters[29]); cap_rtasks entries
	i32* slow_list; i32* slow_list2;    // reads whose segments / reports need the HBM arena (counters[12], counters[13])
	// stage 3: segments of the surviving candidates, alignment jobs, run arena
	KbSegX* segx; u32 cap_segx; u32* cseg_off; i32* cseg_n;   // cseg_* indexed like cands (cseg_n < 0: candidate dropped)
	KbJob* jobs; u32 cap_jobs; u32* runs; u32 cap_runs;
	KbPiece* pieces; u32 cap_pieces;    // nw_alignment problems (cursor: counters[24])
	u32* piece_list;                    // KB_NW_CLASSES lists of cap_pieces piece ids by size class (counts: counters[16 + class])
	u32* part_list;                     // cap_jobs ids of the jobs that go through the 8-mer partition first (count: counters[23])
	// stage 4
	KbReport* reports;                  // indexed like cands
	KbReadRes* res;
	KbPairStat* pstat;
	struct kb_extra_s* extra; u32 cap_extra;      // -m: further lines per read (cursor: counters[14]); kb_extra_t of include/kart_b200.h
	u32* cigar; u32 cap_cigar; u32* cig_cursor;   // cursor: counters[2] of the batch, or the chunk-wide cursor of a pipelined chunk
	// per-thread scratch for the report / rescue kernels
	u8* scratch; u64 scratch_per_thread; i32 scratch_threads;
	u8* wscratch; u64 wscratch_per_warp; i32 wscratch_warps;   // arenas of the warp-per-job kernels (k_align_part, k_nw_warp)
	i32 max_rlen;                       // longest read in the batch
	i32 nw_tmax;                        // largest side one thread solves (<= KB_NW_TMAX); larger problems go to the warp wavefront kernel
	i32 nw_warp_below;                  // a column-tile class with fewer problems than this goes to the warp wavefront kernel as well
	i32 rf_cand;                        // k_rescue_fast: filter-passing window positions noted for the probe phase (the rest is probed on the spot)
	i32 fin_local;                      // k_finalize: pairs with few reports are finished on copies in local memory (kb_stage_finalize)
	i32 rf_reuse, rf_batch;             // k_rescue_fast: keep the mate's 8-mer index across consecutive windows of one mate ; windows per ticket
	i32 rf_stride;                      // k_rescue_fast: 3 = every third window position is scanned (kb_rf_scan), 1 = every position
	i32 part_stack, part_raw;           // k_align_part: entries of a job's work stack / of its exact-match run list that are tried in the warp's shared-memory pool first (the rest, and an overflowing run list, live in the HBM arena)
	i32 seg_cap, kmer_cap;
	// counters: [0] seeds cursor [1] cands cursor [2] cigar cursor [3] status bits [4] rescue count [5] max seeds/read
	//           [6] nw calls [7] rescue attempts [8] segx cursor [10] run cursor + [11] job cursor (one u64) [12],[13] slow lists [14] extra-line cursor (-m) [15] heavy candidate items (k_cand_heavy; listed in slow_list2 until the assemble stage reuses it)
	//           [16..22] pieces per size class [23] partition jobs [24] piece cursor [25] fast rescue tickets [26] k_align_part job tickets [27] rescue task cursor [28] slow rescue tickets [29] slow rescue list
	//           [30],[31] cigar cursor before / after the assemble kernels (pipelined chunks)
	//           64-bit: work[0] extension steps, work[1] occ blocks, work[2] LF steps, work[3] NW cells, work[4] NW calls
	u32* counters;
	unsigned long long* work;
};

#define KB_NCOUNTERS 32
#define KB_NW_CLASSES 7  // nw_alignment size classes: longer side <= 8, 16, 24, 32 (one register tile), <= 64, <= 128 (column tiles of 32), else warp wavefront
#define KB_NW_TMAX 128   // largest side a single thread solves

#endif
