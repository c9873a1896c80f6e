// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stage-level C harness around the UNMODIFIED reference objects (hsinnan75/Kart v2.5.6),
// built into oracle/_ref/libkartref.so by oracle/Makefile. It supplies the globals that the
// reference defines in src/main.cpp:8-14 (main.cpp itself is not linked into the .so) and
// exposes the reference's own per-read functions (prototypes: src/structure.h:177-229,
// src/Mapping.cpp:49-175,317-485) through plain C entry points, so that the tests can pin
// our CPU restatement (kart_oracle.cpp) and the CUDA path stage by stage:
//   seeds (BWT_Search / IdentifySeedPairs_*), candidates, NW, k-mer partition, normal pairs,
//   and the full per-pair / per-read mapping result.
// Dumps are line-oriented text (format documented in oracle/kart_oracle.h) so that the same
// parser compares reference, restatement and GPU output.
#include "structure.h"
#include <string>

// ---- globals normally owned by src/main.cpp -------------------------------------------------
bwt_t *Refbwt;
bwaidx_t *RefIdx;
const char* VersionStr = "2.5.6";
vector<string> ReadFileNameVec1, ReadFileNameVec2;
char *RefSequence, *IndexFileName, *OutputFileName;
int iThreadNum, MaxInsertSize, MaxGaps, MinSeedLength, OutputFileFormat;
bool bDebugMode, bPairEnd, bPacBioData, bMultiHit, gzCompressed, FastQFormat, bSilent;

// ---- reference functions that structure.h does not declare (all have external linkage) ------
extern void RemoveRedundantCandidates(vector<AlignmentCandidate_t>& AlignmentVec);          // Mapping.cpp:317
extern void RemoveUnMatedAlignmentCandidates(vector<AlignmentCandidate_t>&, vector<AlignmentCandidate_t>&); // :402
extern void CheckPairedFinalAlignments(ReadItem_t& read1, ReadItem_t& read2);               // :429
extern void SetSingleAlignmentFlag(ReadItem_t& read);                                       // :49
extern void SetPairedAlignmentFlag(ReadItem_t& read1, ReadItem_t& read2);                   // :73
extern void EvaluateMAPQ(ReadItem_t& read);                                                 // :160
extern void EnCodeReadSeq(int rlen, char* seq, uint8_t* EncodeSeq);                         // :482
extern bool CheckLocalAlignmentQuality(string& aln1, string& aln2);                         // tools.cpp:255
extern bwtint_t bwt_sa(bwtint_t k);                                                         // bwt_search.cpp:128
extern void bwt_occ4(const bwt_t *bwt, bwtint_t k, bwtint_t cnt[4]);                         // bwt_search.cpp:68

static long emit(const std::string& s, char* out, long cap)
{
	if (out != NULL && (long)s.size() < cap) { memcpy(out, s.data(), s.size()); out[s.size()] = '\0'; }
	return (long)s.size();
}

static void put(std::string& s, const char* fmt, ...)
{
	va_list ap, ap2; va_start(ap, fmt); va_copy(ap2, ap);
	int n = vsnprintf(NULL, 0, fmt, ap); va_end(ap);
	std::vector<char> buf((size_t)n + 1); vsnprintf(buf.data(), buf.size(), fmt, ap2); va_end(ap2);
	s.append(buf.data(), n);
}

static void dump_seeds(std::string& s, const vector<SeedPair_t>& v)
{
	for (size_t i = 0; i < v.size(); i++)
		put(s, "S %d %d %d %lld %d\n", v[i].rPos, v[i].rLen, v[i].gLen, (long long)v[i].gPos, v[i].bSimple ? 1 : 0);
}

static void dump_cands(std::string& s, const vector<AlignmentCandidate_t>& v)
{
	for (size_t i = 0; i < v.size(); i++)
	{
		put(s, "C %d %lld %d %d\n", v[i].Score, (long long)v[i].PosDiff, v[i].PairedAlnCanIdx, (int)v[i].SeedVec.size());
		dump_seeds(s, v[i].SeedVec);
	}
}

static void dump_read(std::string& s, const ReadItem_t& r)
{
	put(s, "R %d %d %d %d %d\n", r.score, r.sub_score, r.mapq, r.CanNum, r.iBestAlnCanIdx);
	for (int i = 0; i < r.CanNum; i++)
	{
		const AlignmentReport_t& a = r.AlnReportArr[i];
		bool flagged = (r.score == 0 && i == 0) || (r.score > 0 && i == r.iBestAlnCanIdx);
		put(s, "A %d %d %d", i, a.AlnScore, a.PairedAlnCanIdx);
		if (flagged) put(s, " F%d", a.SamFlag);
		if (a.AlnScore > 0) put(s, " %d %d %lld %s", a.coor.bDir ? 1 : 0, a.coor.ChromosomeIdx, (long long)a.coor.gPos, a.coor.CIGAR.c_str());
		s += "\n";
	}
}

extern "C" {

// Load index + reference exactly as main.cpp:192-208 / Mapping.cpp:645 do.
int kref_load(const char* prefix, int pacbio, int max_gaps, int multihit, int threads)
{
	MaxGaps = max_gaps; iThreadNum = threads > 0 ? threads : 1; bPairEnd = false; bDebugMode = false; MaxInsertSize = 1500;
	bPacBioData = pacbio != 0; bMultiHit = multihit != 0; bSilent = true; FastQFormat = true; OutputFileFormat = 0;
	OutputFileName = (char*)"output.sam"; IndexFileName = (char*)prefix;
	if (!CheckBWAIndexFiles(prefix)) return -1;
	fflush(stdout); int so = dup(1); FILE* nul = fopen("/dev/null", "w"); dup2(fileno(nul), 1);   // the loader chats on stdout
	RefIdx = bwa_idx_load(prefix);
	if (RefIdx != 0) { Refbwt = RefIdx->bwt; RestoreReferenceInfo(); }
	fflush(stdout); dup2(so, 1); close(so); fclose(nul);
	if (RefIdx == 0) return -2;
	for (MinSeedLength = 13; MinSeedLength < 16; MinSeedLength++) if (TwoGenomeSize < pow(4, MinSeedLength)) break;
	return 0;
}

void kref_set_mode(int pacbio, int max_gaps, int multihit) { bPacBioData = pacbio != 0; MaxGaps = max_gaps; bMultiHit = multihit != 0; }
int kref_min_seed_len() { return MinSeedLength; }
long long kref_genome_size() { return (long long)GenomeSize; }
const char* kref_refseq() { return RefSequence; }

void kref_occ4(unsigned long long k, unsigned long long* cnt) { bwtint_t c[4]; bwt_occ4(Refbwt, (bwtint_t)k, c); for (int i = 0; i < 4; i++) cnt[i] = c[i]; }
unsigned long long kref_sa(unsigned long long k) { return (unsigned long long)bwt_sa((bwtint_t)k); }

// BWT_Search on nt4 codes; locs must hold 50 entries.
void kref_bwt_search(const uint8_t* codes, int start, int stop, int* len, int* freq, unsigned long long* locs)
{
	bwtSearchResult_t r = BWT_Search((uint8_t*)codes, start, stop);
	*len = r.len; *freq = r.freq;
	for (int i = 0; i < r.freq; i++) locs[i] = r.LocArr[i];
	if (r.freq > 0) delete[] r.LocArr;
}

long kref_seeds(const char* seq, int rlen, int sensitive, char* out, long cap)
{
	std::string s; uint8_t* enc = new uint8_t[rlen + 1]; EnCodeReadSeq(rlen, (char*)seq, enc);
	vector<SeedPair_t> v = sensitive ? IdentifySeedPairs_SensitiveMode(rlen, enc) : IdentifySeedPairs_FastMode(rlen, enc);
	delete[] enc; dump_seeds(s, v);
	return emit(s, out, cap);
}

long kref_candidates(const char* seq, int rlen, char* out, long cap)
{
	std::string s; uint8_t* enc = new uint8_t[rlen + 1]; EnCodeReadSeq(rlen, (char*)seq, enc);
	vector<AlignmentCandidate_t> c;
	if (bPacBioData) c = GenerateAlignmentCandidateForPacBioSeq(rlen, IdentifySeedPairs_SensitiveMode(rlen, enc));
	else c = GenerateAlignmentCandidateForIlluminaSeq(rlen, IdentifySeedPairs_FastMode(rlen, enc));
	delete[] enc; dump_cands(s, c);
	return emit(s, out, cap);
}

// nw_alignment on raw strings; o1/o2 need m+n+1 bytes each. Returns aligned length.
long kref_nw(const char* s1, int m, const char* s2, int n, char* o1, char* o2)
{
	string a(s1, m), b(s2, n);
	nw_alignment(m, a, n, b);
	memcpy(o1, a.c_str(), a.size() + 1); memcpy(o2, b.c_str(), b.size() + 1);
	return (long)a.size();
}

// GenerateSimplePairsFromFragmentPair (KmerAnalysis.cpp:164) followed, if do_normal, by IdentifyNormalPairs(len1,len2,..)
long kref_fragment_pairs(int max_dist, const char* f1, int len1, const char* f2, int len2, int do_normal, char* out, long cap)
{
	std::string s; string a(f1, len1), b(f2, len2);
	vector<SeedPair_t> v = GenerateSimplePairsFromFragmentPair(max_dist, len1, (char*)a.c_str(), len2, (char*)b.c_str());
	if (do_normal && v.size() > 0) IdentifyNormalPairs(len1, len2, v);
	dump_seeds(s, v);
	return emit(s, out, cap);
}

// IdentifyNormalPairs on a caller-supplied simple-pair list (n entries of rPos,rLen,gPos; gLen=rLen).
long kref_normal_pairs(int rlen, int glen, int n, const int* rpos, const int* rl, const long long* gpos, char* out, long cap)
{
	std::string s; vector<SeedPair_t> v(n);
	for (int i = 0; i < n; i++) { v[i].bSimple = true; v[i].rPos = rpos[i]; v[i].rLen = v[i].gLen = rl[i]; v[i].gPos = gpos[i]; v[i].PosDiff = gpos[i] - rpos[i]; }
	IdentifyNormalPairs(rlen, glen, v);
	dump_seeds(s, v);
	return emit(s, out, cap);
}

// kind: 0 = ProcessNormalSequencePair, 1 = Head, 2 = Tail (tools.cpp:225,292,344).
// Output: "P score rPos rLen gPos gLen\n" then one "O len op" line per cigar element pushed.
long kref_process_pair(int kind, const char* seq, int rpos, int rlen, long long gpos, int glen, char* out, long cap)
{
	std::string s; SeedPair_t sp; sp.bSimple = false; sp.rPos = rpos; sp.rLen = rlen; sp.gPos = gpos; sp.gLen = glen; sp.PosDiff = gpos - rpos;
	vector<pair<int, char> > cig; int score;
	if (kind == 0) score = ProcessNormalSequencePair((char*)seq, sp, cig);
	else if (kind == 1) score = ProcessHeadSequencePair((char*)seq, sp, cig);
	else score = ProcessTailSequencePair((char*)seq, sp, cig);
	put(s, "P %d %d %d %lld %d\n", score, sp.rPos, sp.rLen, (long long)sp.gPos, sp.gLen);
	for (size_t i = 0; i < cig.size(); i++) put(s, "O %d %c\n", cig[i].first, cig[i].second);
	return emit(s, out, cap);
}

static void free_read(ReadItem_t& r) { if (r.CanNum > 0) delete[] r.AlnReportArr; }

// One read through the single-end / pacbio branch of ReadMapping (Mapping.cpp:513-529, 580-596).
long kref_map_single(const char* seq, int rlen, char* out, long cap)
{
	std::string s; ReadItem_t r; memset(&r, 0, sizeof(r));
	string sq(seq, rlen); r.rlen = rlen; r.seq = (char*)sq.c_str(); r.header = (char*)"r"; r.qual = NULL;
	uint8_t* enc = new uint8_t[rlen + 1]; EnCodeReadSeq(rlen, r.seq, enc);
	vector<SeedPair_t> sv; vector<AlignmentCandidate_t> av;
	if (bPacBioData) { sv = IdentifySeedPairs_SensitiveMode(rlen, enc); av = GenerateAlignmentCandidateForPacBioSeq(rlen, sv); }
	else { sv = IdentifySeedPairs_FastMode(rlen, enc); av = GenerateAlignmentCandidateForIlluminaSeq(rlen, sv); }
	delete[] enc;
	RemoveRedundantCandidates(av);
	GenMappingReport(true, r, av);
	SetSingleAlignmentFlag(r); EvaluateMAPQ(r);
	dump_read(s, r); free_read(r);
	return emit(s, out, cap);
}

// One pair through the paired branch (Mapping.cpp:542-578). seq2 must already be reverse-complemented
// (GetData.cpp:125-135). Also reports the insert-size statistic contribution of OutputPairedAlignments (:206-213):
// "P counted absdist".  stage_dump != 0 additionally emits the candidate lists right before GenMappingReport.
long kref_map_pair(const char* seq1, int l1, const char* seq2, int l2, int est, int stage_dump, char* out, long cap)
{
	std::string s; ReadItem_t r1, r2; memset(&r1, 0, sizeof(r1)); memset(&r2, 0, sizeof(r2));
	string a(seq1, l1), b(seq2, l2);
	r1.rlen = l1; r1.seq = (char*)a.c_str(); r1.header = (char*)"r1"; r2.rlen = l2; r2.seq = (char*)b.c_str(); r2.header = (char*)"r2";
	uint8_t* enc = new uint8_t[(l1 > l2 ? l1 : l2) + 1];
	EnCodeReadSeq(l1, r1.seq, enc); vector<SeedPair_t> sv1 = IdentifySeedPairs_FastMode(l1, enc);
	vector<AlignmentCandidate_t> av1 = GenerateAlignmentCandidateForIlluminaSeq(l1, sv1);
	EnCodeReadSeq(l2, r2.seq, enc); vector<SeedPair_t> sv2 = IdentifySeedPairs_FastMode(l2, enc);   // reference passes l1 here (Mapping.cpp:550); equal lengths in all tests
	vector<AlignmentCandidate_t> av2 = GenerateAlignmentCandidateForIlluminaSeq(l2, sv2);
	delete[] enc;
	bool paired = CheckPairedAlignmentCandidates(est, av1, av2);
	int rescued = 0;
	if (!paired) { paired = RescueUnpairedAlignment(est, r1, r2, av1, av2); rescued = paired ? 1 : 0; }
	if (paired) RemoveUnMatedAlignmentCandidates(av1, av2);
	RemoveRedundantCandidates(av1); RemoveRedundantCandidates(av2);
	if (stage_dump) { put(s, "X %d %d\n", paired ? 1 : 0, rescued); s += "V1\n"; dump_cands(s, av1); s += "V2\n"; dump_cands(s, av2); }
	GenMappingReport(true, r1, av1); GenMappingReport(false, r2, av2);
	CheckPairedFinalAlignments(r1, r2);
	SetPairedAlignmentFlag(r1, r2);
	EvaluateMAPQ(r1); EvaluateMAPQ(r2);
	dump_read(s, r1); dump_read(s, r2);
	int counted = 0, ad = 0;
	if (r1.score > 0)
	{
		int i = r1.iBestAlnCanIdx, j;
		if (r1.AlnReportArr[i].AlnScore > 0 && (j = r1.AlnReportArr[i].PairedAlnCanIdx) != -1 && r2.AlnReportArr[j].AlnScore > 0)
		{
			int dist = (int)(r2.AlnReportArr[j].coor.gPos - r1.AlnReportArr[i].coor.gPos + (r1.AlnReportArr[i].coor.bDir ? r2.rlen : 0 - r1.rlen));
			counted = 1; ad = abs(dist);
		}
	}
	put(s, "P %d %d\n", counted, ad);
	free_read(r1); free_read(r2);
	return emit(s, out, cap);
}

} // extern "C"
