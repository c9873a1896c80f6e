/* kart_b200 -- C ABI of the B200 (sm_100a) implementation of Kart's per-read hot path.
 *
 * Kart (hsinnan75/Kart v2.5.6) is a monolithic executable without a plugin or FFI layer; the seam this library
 * replaces is the per-chunk body of ReadMapping(), reference src/Mapping.cpp:513-596: everything between
 * "chunk read" (GetNextChunk, :506-510) and "format SAM" (OutputPaired/SingledAlignments, :597-599).
 * The reference functions behind that seam (prototypes in src/structure.h:177-229):
 *   BWT_Search (structure.h:178), IdentifySeedPairs_FastMode/_SensitiveMode (:189,:191),
 *   GenerateAlignmentCandidateForIlluminaSeq/ForPacBioSeq (:193,:194), CheckPairedAlignmentCandidates (:182),
 *   RescueUnpairedAlignment (:198), GenMappingReport (:192), nw_alignment (:229), and from src/Mapping.cpp
 *   RemoveRedundantCandidates (:317), RemoveUnMatedAlignmentCandidates (:402), CheckPairedFinalAlignments (:429),
 *   SetPairedAlignmentFlag (:73), SetSingleAlignmentFlag (:49), EvaluateMAPQ (:160).
 *
 * Plain C types only; no CUDA, torch or C++ types cross this boundary. All functions return 0 on success or a
 * negative KB_E* code (kb_strerror). There is no CPU fallback: without a CUDA device every compute entry point fails.
 * Threading: one kb_ctx_t per device and per host thread (mirrors one ReadMapping pthread, src/Mapping.cpp:716).
 */
#ifndef KART_B200_H
#define KART_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb_ctx kb_ctx_t;

enum { KB_OK = 0, KB_ENODEV = -1, KB_ECUDA = -2, KB_EINVAL = -3, KB_ENOMEM = -4, KB_ENOINDEX = -5, KB_ECAPACITY = -6, KB_EOVERFLOW = -7, KB_ESTATE = -8 };

/* Host view of the BWA-format index exactly as the reference loads it (src/bwt_index.cpp:16-36,103-122,148-259).
 * Pointers are borrowed for the duration of kb_upload_index only. */
typedef struct {
	uint64_t primary;            /* .bwt header word 0                                    (bwt_index.cpp:114) */
	uint64_t L2[5];              /* L2[0] = 0, L2[1..4] = .bwt header words 1..4           (:115)              */
	uint64_t seq_len;            /* = L2[4]                                                (:117)              */
	const uint32_t* bwt;         /* .bwt payload: interleaved Occ/BWT, 16 words / 128 rows (BWT_Index/bwtindex.c:53) */
	uint64_t bwt_words;
	const uint64_t* sa;          /* sampled SA with sa[0] = ~0 prepended, n_sa entries     (bwt_index.cpp:30-34) */
	uint64_t n_sa;
	int32_t sa_intv;             /* 32                                                                       */
	const uint8_t* pac;          /* forward strand, 2 bit/base, l_pac/4+1 bytes            (bwt_index.cpp:156,240) */
	int64_t l_pac;               /* GenomeSize                                                                */
	int32_t n_chr;
	const int64_t* chr_len;      /* bntann1_t::len per sequence, in .ann order             (:244-251)         */
} kb_index_host_t;

/* Globals of the reference that steer the hot path (src/structure.h:157-170, src/main.cpp:92-104). */
typedef struct {
	int32_t min_seed_len;        /* MinSeedLength; <= 0: derive from 2*l_pac as Mapping.cpp:645 does */
	int32_t max_gaps;            /* -g   [5]    */
	int32_t max_insert;          /* MaxInsertSize [1500] */
	int32_t pacbio;              /* -pacbio     */
	int32_t paired;              /* reads are (mate 1, reverse-complemented mate 2) pairs: ReadMapping's bPairEnd branch (:531) */
	int32_t multihit;            /* -m          */
} kb_params_t;

/* One chunk of reads, structure-of-arrays. seq holds the raw read characters back to back; for paired input read 2i+1
 * is the mate of read 2i and must already be reverse-complemented the way GetNextChunk does (src/GetData.cpp:125-135). */
typedef struct {
	int32_t n_reads;
	const uint8_t* seq;
	const uint64_t* seq_off;     /* n_reads + 1 offsets into seq */
} kb_reads_t;

/* The same chunk in the form the device works on, for callers that pack while parsing (FASTQ text is 8 bits per base; over PCIe
 * that is the longest pole of kb_map_chunk): 2-bit codes (nst_nt4_table, src/BWT_Index/bntseq.c:40: A/a 0, C/c 1, G/g 2, T/t 3), 32
 * bases per 64-bit word, base i of a word at bits 62-2i, zero where the character is no base and behind the end of a read. Read r
 * starts at word (seq_off[r] >> 5) + r and takes ceil(len / 32) words; n_words = (seq_off[n_reads] >> 5) + n_reads + 1.
 * Every character that is not one of the upper-case letters A C G T is listed in exc, ascending:
 * (read index << 32) | (position in the read << 8) | character -- the device restores the exact characters from it, so results
 * are identical to kb_map_chunk on the text (N, lower case and IUPAC codes keep their reference semantics). kb_pack_reads() makes
 * this form from a kb_reads_t. */
typedef struct {
	int32_t n_reads;
	const uint64_t* code;
	uint64_t n_words;
	const uint64_t* seq_off;     /* n_reads + 1 character offsets, as in kb_reads_t */
	const uint64_t* exc;
	uint64_t n_exc;
} kb_reads_packed_t;

/* What OutputPairedAlignments / OutputSingledAlignments (src/Mapping.cpp:177-315) need to print one read's primary line. */
typedef struct {
	int64_t pos;                 /* 1-based leftmost coordinate (Coordinate_t::gPos)                */
	int64_t mate_pos;            /* RNEXT/PNEXT coordinate, or -1 when the line carries "*\t0\t0"   */
	int32_t kind;                /* 0: unmapped line, 1: mapped line, 2: no line (score>0 but best report cleared) */
	int32_t flag;                /* SAM flag                                                        */
	int32_t chr;                 /* index into the .ann sequence list                               */
	int32_t mapq;
	int32_t score, sub_score;    /* AS / XS ; NM = rlen - score                                     */
	int32_t tlen;
	int32_t fwd;                 /* Coordinate_t::bDir                                              */
	uint32_t cig_off;            /* first op in the cigar array                                     */
	int32_t cig_len;             /* ops: len << 4 | op with BAM op numbers (M=0,I=1,D=2,S=4)        */
} kb_aln_t;

/* -m (bMultiHit): a read prints one line per report from iBestAlnCanIdx on (src/Mapping.cpp:194-223,242-263,289-308). kb_aln_t
 * carries the first of them; every further line is one kb_extra_t. rank orders the extra lines of one read. */
typedef struct kb_extra_s {
	uint32_t read;               /* index of the read in the chunk                                  */
	uint32_t rank;               /* 0, 1, ... in AlnReportArr order                                 */
	kb_aln_t aln;                /* kind is always 1; cig_off points into the chunk's cigar array   */
} kb_extra_t;

/* Per pair: contribution to the running insert-size statistic (iPaired/iDistance, src/Mapping.cpp:206-213) and the
 * closed interval of EstDistance values for which this pair's result is provably unchanged (host-side recurrence,
 * src/Mapping.cpp:533-540). */
typedef struct { int32_t counted, absdist, est_lo, est_hi; } kb_pair_stat_t;

/* Caller-owned result buffers (pinned memory makes the copies asynchronous, pageable works too). */
typedef struct {
	kb_aln_t* aln;               /* n_reads entries                                    */
	kb_pair_stat_t* pairs;       /* n_reads/2 entries, may be NULL when not paired     */
	uint32_t* cigar;             /* cap_cigar entries                                  */
	uint32_t cap_cigar;
	uint32_t n_cigar;            /* out: entries used (also set when KB_ECAPACITY is returned) */
} kb_results_t;

int  kb_device_count(void);                 /* visible CUDA devices (0 without a driver or a device) */
int  kb_init(int device, kb_ctx_t** out);
void kb_destroy(kb_ctx_t* ctx);
const char* kb_strerror(int code);
const char* kb_last_error(kb_ctx_t* ctx);

/* Copies the index to the device and re-blocks it (see kart_b200/csrc/kb_types.h). expand_sa = 1 additionally expands
 * the sampled SA into a full SA on the device (8 bytes per BWT row; values identical by construction, the .sa file format is
 * untouched): locates become one load, and a search that is down to one row finishes by comparing the read with the text.
 * expand_sa = 2 does so when the device has the memory to spare (the full SA plus 48 GB for batches), 0 never. */
int  kb_upload_index(kb_ctx_t* ctx, const kb_index_host_t* idx, int expand_sa);
/* Gives `dst` (another device) the index `src` holds, device to device (NVLink between peers): no host copy, no table build,
 * no SA expansion. The reference shares one read-only index between its worker threads (Refbwt / RefSequence, src/Mapping.cpp:716);
 * on several GPUs every device needs its own replica. */
int  kb_clone_index(kb_ctx_t* dst, kb_ctx_t* src);
int  kb_set_params(kb_ctx_t* ctx, const kb_params_t* p);
int  kb_get_min_seed_len(kb_ctx_t* ctx);

/* The drop-in call: host buffers in, host buffers out (H2D, all kernels, D2H). est holds one EstDistance per pair
 * (n_reads/2 values) when paired, else it may be NULL. */
int  kb_map_chunk(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est, kb_results_t* out);

/* kb_map_chunk / kb_stage_reads on packed reads (48 instead of 160 bytes per 150-bp read over PCIe). */
int  kb_map_chunk_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est, kb_results_t* out);
int  kb_stage_reads_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est);
/* Chunks in flight: the same call split in two, for callers that have the next chunk ready while one is being mapped (a mapper
 * streaming a read file). Every chunk runs as one batch on a stream of its own, so the H2D copy of chunk k+1 and the D2H copy of
 * chunk k-1 run under the kernels of chunk k; at most 3 chunks may be in flight. begin returns immediately with a ticket; the read
 * and result buffers must stay valid and untouched until end(ticket) returns, which delivers out->n_cigar and the cigar elements.
 * Chunks may be ended in any order. Not available with -m. */
int  kb_map_chunk_begin(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est, kb_results_t* out, int* ticket);
int  kb_map_chunk_begin_packed(kb_ctx_t* ctx, const kb_reads_packed_t* in, const int32_t* est, kb_results_t* out, int* ticket);
int  kb_map_chunk_end(kb_ctx_t* ctx, int ticket);
/* Host helper: packs `in` into caller-owned buffers (code: kb_packed_words(in) words; exc: cap_exc entries) on `threads` host
 * threads and fills *out (which borrows in->seq_off, code and exc). KB_ECAPACITY when exc is too small (out->n_exc then holds
 * the count needed). Needs no device. */
uint64_t kb_packed_words(const kb_reads_t* in);
int  kb_pack_reads(const kb_reads_t* in, uint64_t* code, uint64_t* exc, uint64_t cap_exc, int threads, kb_reads_packed_t* out);

/* The same pipeline in three steps, for callers that keep a batch resident in HBM (bench.py's device-timed `value`). */
int  kb_stage_reads(kb_ctx_t* ctx, const kb_reads_t* in, const int32_t* est);   /* H2D */
int  kb_run(kb_ctx_t* ctx);                                                      /* kernels only, returns after they finish */
int  kb_fetch_results(kb_ctx_t* ctx, kb_results_t* out);                         /* D2H */

/* -m only: the extra lines of the last kb_map_chunk / kb_run, sorted by (read, rank). *n receives the number of entries the
 * chunk has; when that exceeds cap nothing is copied and KB_ECAPACITY is returned (call again with a larger buffer).
 * Chunks are not pipelined while multihit is set. */
int  kb_fetch_extra(kb_ctx_t* ctx, kb_extra_t* out, uint32_t cap, uint32_t* n);

/* Page-locked host memory for the caller-owned buffers (reads in, results out): with it the copies of kb_map_chunk are
 * asynchronous DMA and overlap the kernels; pageable buffers work but are staged by the driver. NULL when no device /
 * no memory. Usable from any thread, before or after kb_init. */
void* kb_host_alloc(uint64_t bytes);
void  kb_host_free(void* p);
/* Page-locks memory the caller allocated itself (page-aligned is best), e.g. buffers filled while the device was still being
 * initialised. Returns KB_OK or KB_ECUDA; a buffer that could not be registered still works, as pageable memory. */
int   kb_host_register(void* p, uint64_t bytes);
void  kb_host_unregister(void* p);

/* Instrumentation. kb_stage_ms: device time (CUDA events on the context's stream) of each kernel of the last kb_run:
 * [0] fm_seed [1] sa_locate [2] cand_pair [3] rescue [4] segments [5] align [6] assemble [7] finalize [8] whole run
 * [9] the nw_alignment solver kernels alone (part of [5]).
 * Returns entries written.
 * kb_work: algorithmic work of the last run: [0] extension steps [1] Occ blocks touched (32-byte sectors)
 * [2] LF steps [3] NW cells [4] seeds [5] NW calls [6] rescue attempts [7] kernels launched. */
int  kb_stage_ms(kb_ctx_t* ctx, float* ms, int n);
int  kb_work(kb_ctx_t* ctx, uint64_t* w, int n);
void* kb_cuda_stream(kb_ctx_t* ctx);

/* Test hook: copy an internal device array of the last run to the host. what: 0 n_seeds(i32) 1 seed_off(u32) 2 segs(KbSeg)
 * 3 n_cands(i32) 4 cand_off(u32) 5 cands(KbCand) 6 reports(KbReport) 7 res(KbReadRes) 8 cigar(u32) 9 counters(u32[32]) 10 jobs(KbJob)
 * 11 hits(KbHit: the searches that yield seeds, [read][max_hits]) 12 n_hits(i32) 13 max_hits(i32).
 * Returns bytes copied (<= bytes) or a negative error. */
int64_t kb_debug_fetch(kb_ctx_t* ctx, int what, void* dst, uint64_t bytes);

/* Test hook, stage level: runs caller-chosen fragment pairs of the staged reads (kb_stage_reads) through the device code behind
 * Process{Normal,Head,Tail}SequencePair (src/tools.cpp:225,292,344): the quick tests, the 8-mer partition
 * (GenerateNormalPairAlignment, tools.cpp:142; k_align_part), nw_alignment (src/nw_alignment.cpp:18; k_nw_tile<*>, k_nw_warp),
 * k_align_gather and the per-segment cigar rules of GenMappingReport (src/AlignmentCandidates.cpp:648-723).
 * mode: 0 middle pair, 1 head, 2 tail (as the reference classifies them), 3 straight to GenerateNormalPairAlignment (no quick
 * test), 4 straight to ONE nw_alignment call whatever the size. ops receives every fragment's cigar elements (len << 4 | op)
 * at out[i].ops_off, out[i].n_ops of them; cap_ops must be at least the sum of rlen + glen + 4. */
typedef struct { uint32_t read; int32_t rpos, rlen, glen, mode, pad; int64_t gpos; } kb_dbg_frag_t;
typedef struct { int32_t info, aux, score, n_ops; uint32_t ops_off; int32_t nruns, ident, aligned; int64_t g_first, g_end; } kb_dbg_frag_out_t;
int kb_debug_align(kb_ctx_t* ctx, const kb_dbg_frag_t* specs, int n, kb_dbg_frag_out_t* out, uint32_t* ops, uint32_t cap_ops);

/* Index construction on the device (`kart index -gpu`): the contents of the .bwt and .sa files of the 2G text (forward strand +
 * reverse complement) from the packed forward strand, i.e. the BWT / Occ / SA part of the reference's bwa_idx_build
 * (src/BWT_Index/bwtindex.c:107-142: bwt_pac2bwt -> bwt_gen.c:1601, bwt_bwtupdate_core :53-75, bwt_cal_sa bwt.c:101-123); byte-identical
 * output. pac: l_pac/4 + 1 bytes, 2 bit per base, MSB first (ambiguous bases already replaced, bntseq.c:144).
 * out->bwt: the .bwt payload behind its 40-byte header (bwt_words 32-bit words); out->sa: the n_sa - 1 samples sa[1..] (the .sa payload
 * behind its 56-byte header). Both are page-locked host arrays owned by the library until kb_index_free(). */
typedef struct { uint64_t primary; uint64_t L2[5]; uint64_t seq_len; uint32_t* bwt; uint64_t bwt_words; uint64_t* sa; uint64_t n_sa; } kb_built_index_t;
int  kb_index_build(int device, const uint8_t* pac, int64_t l_pac, kb_built_index_t* out);
void kb_index_free(kb_built_index_t* b);
const char* kb_index_build_error(void);

#ifdef __cplusplus
}
#endif
#endif
