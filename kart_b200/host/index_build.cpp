// `kart index ref.fa prefix`: BWA-format index files, byte-identical to what the reference's builder writes
// (reference src/BWT_Index/bwtindex.c:77-149 bwa_idx_build; bntseq.c:160-215 bns_fasta2bntseq, :59-89 bns_dump;
// bwtindex.c:53-75 bwt_bwtupdate_core; bwt.c:101-123 bwt_cal_sa, :174-196 bwt_dump_bwt / bwt_dump_sa).
//
// Freshly written and organised differently: the reference builds the BWT incrementally (BWT-SW, bwt_gen.c), rewrites the
// .bwt file twice and derives the sampled SA by walking the whole BWT backwards. The files are canonical functions of the text,
// so here the suffix array of the 2G text (forward strand + reverse complement) is built directly -- suffixes are bucketed
// by their first 12 bases with a counting sort and every bucket is finished with a comparison sort over the 2-bit packed
// text, 32 bases per step, buckets spread over threads -- and the BWT, the interleaved Occ blocks and the SA samples are read
// off it. Suffix positions are 32-bit when 2G + 2 < 2^32 (genomes up to 2.1 Gbp) and 64-bit beyond (KART_INDEX_64=1 forces
// the 64-bit instantiation; the tests run both on the same genomes).
#include "kart_host.h"
#include <algorithm>
#include <atomic>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <time.h>

namespace {

struct Ann { std::string name, anno; long long offset; int len, n_ambs; };
struct Amb { long long offset; int len; char amb; };

inline int nt4(unsigned char c)   // nst_nt4_table, bntseq.c:40-57
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

// FASTA/FASTQ entry reader with kseq's rules (kseq.h:180-215): name up to the first white space, comment = rest of the header
// line, sequence = every character of the following lines except '\n' (a '\r' ending a line of more than one character is
// dropped), up to the next line starting with '>', '@' or '+'.
struct FastaReader
{
	gzFile fp = nullptr; std::vector<unsigned char> buf; size_t at = 0, end = 0; bool eof = false; int last_char = 0;
	int getc() { if (at >= end) { if (eof) return -1; int n = gzread(fp, buf.data(), (unsigned)buf.size()); if (n <= 0) { eof = true; return -1; } at = 0; end = (size_t)n; } return buf[at++]; }
	// reads up to (and consuming) a delimiter: mode 0 = white space, 1 = end of line. Returns false when nothing was left to read.
	bool until(int mode, std::string& s, int* delim, bool append)
	{
		bool any = false; if (delim) *delim = 0; if (!append) s.clear();
		while (true)
		{
			if (at >= end) { if (eof) break; int n = gzread(fp, buf.data(), (unsigned)buf.size()); if (n <= 0) { eof = true; break; } at = 0; end = (size_t)n; }
			size_t i = at;
			if (mode == 1) { while (i < end && buf[i] != '\n') i++; } else { while (i < end && !isspace(buf[i])) i++; }
			any = true; s.append((const char*)buf.data() + at, i - at); size_t stop = i; at = i + 1;
			if (stop < end) { if (delim) *delim = buf[stop]; break; }
		}
		if (!any && eof) return false;
		if (mode == 1 && s.size() > 1 && s.back() == '\r') s.pop_back();
		return true;
	}
	bool next(std::string& name, std::string& comment, std::string& seq)
	{
		int c;
		if (last_char == 0) { while ((c = getc()) != -1 && c != '>' && c != '@') {} if (c == -1) return false; last_char = c; }
		comment.clear(); seq.clear();
		if (!until(0, name, &c, false)) return false;
		if (c != '\n') until(1, comment, nullptr, false);
		while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') { if (c == '\n') continue; seq.push_back((char)c); until(1, seq, nullptr, true); }
		if (c == '>' || c == '@') last_char = c;
		if (c != '+') return true;
		std::string qual; while ((c = getc()) != -1 && c != '\n') {}
		if (c == -1) return true;
		while (until(1, qual, nullptr, true) && qual.size() < seq.size()) {}
		last_char = 0;
		return true;
	}
};

// forward strand as 2-bit codes; ambiguous characters become lrand48()&3 with the fixed seed 11 (bntseq.c:144,173), runs of one
// and the same ambiguous character are one hole (:124-139)
bool read_reference(const char* fa, std::vector<uint8_t>& codes, std::vector<Ann>& anns, std::vector<Amb>& ambs)
{
	FastaReader rd; rd.fp = gzopen(fa, "r"); if (!rd.fp) return false; rd.buf.resize(1 << 20);
	srand48(11);
	std::string name, comment, seq;
	while (rd.next(name, comment, seq))
	{
		Ann a; a.name = name; a.anno = comment.empty() ? "(null)" : comment; a.len = (int)seq.size(); a.n_ambs = 0;
		a.offset = anns.empty() ? 0 : anns.back().offset + anns.back().len;
		int lasts = 0;
		for (size_t i = 0; i < seq.size(); i++)
		{
			int c = nt4((unsigned char)seq[i]);
			if (c >= 4)
			{
				if (lasts == seq[i]) ambs.back().len++;
				else { Amb h; h.len = 1; h.offset = a.offset + (long long)i; h.amb = seq[i]; ambs.push_back(h); a.n_ambs++; }
			}
			lasts = seq[i];
			if (c >= 4) c = (int)(lrand48() & 3);
			codes.push_back((uint8_t)c);
		}
		anns.push_back(a);
	}
	gzclose(rd.fp);
	return true;
}

void write_or_die(FILE* fp, const void* p, size_t n) { if (n && fwrite(p, 1, n, fp) != n) { fprintf(stderr, "Error! write failed\n"); exit(1); } }

// 2G text, 2 bits per base, 32 bases per big-endian word (base i of a word at bits 62-2i), zero padded
struct PackedText
{
	std::vector<uint64_t> w; uint64_t n = 0;
	uint64_t win(uint64_t p) const   // 32 bases starting at p (zeros beyond the end)
	{
		const uint64_t i = p >> 5; const int s = (int)(p & 31) * 2;
		return s ? (w[i] << s) | (w[i + 1] >> (64 - s)) : w[i];
	}
	// suffix order with the end of the text smaller than every base
	template <class IDX> bool less(IDX a, IDX b) const
	{
		if (a == b) return false;
		uint64_t x = a, y = b;
		while (x < n && y < n)
		{
			const uint64_t u = win(x), v = win(y);
			if (u != v)
			{
				const uint64_t d = (uint64_t)__builtin_clzll(u ^ v) >> 1;   // first differing base
				if (x + d >= n || y + d >= n) break;                          // the difference lies in the padding of one of them
				return u < v;
			}
			x += 32; y += 32;
		}
		return a > b;   // one suffix is a prefix of the other: the shorter (larger start) comes first
	}
};

}   // namespace

static int write_pac_ann_amb(const std::string& prefix, const std::vector<uint8_t>& fwd, const std::vector<Ann>& anns, const std::vector<Amb>& ambs, std::vector<uint8_t>* pac_out);

template <class IDX>
static int build_core(const std::string& prefix, const std::vector<uint8_t>& fwd, const std::vector<Ann>& anns, const std::vector<Amb>& ambs, int threads)
{
	const uint64_t L = fwd.size(), N = 2 * L;
	clock_t t0 = clock();

	// ---- the 2G text and its suffix array ----
	fprintf(stdout, "[bwt_index] Construct BWT for the packed sequence...\n"); fflush(stdout);
	time_t w0 = time(NULL);
	auto base = [&](uint64_t i) -> unsigned { return i < L ? fwd[i] : 3u - fwd[N - 1 - i]; };   // bntseq.c:190-191
	PackedText T; T.n = N; T.w.assign(N / 32 + 3, 0);
	const int nt = std::max(1, threads);
	parallel_for(nt, (size_t)((N + 31) / 32), [&](int, size_t w0_, size_t w1_) {
		for (size_t w = w0_; w < w1_; w++) { uint64_t v = 0; const uint64_t e = std::min<uint64_t>(N, 32ull * w + 32); for (uint64_t i = 32ull * w; i < e; i++) v |= (uint64_t)base(i) << (62 - 2 * (i & 31)); T.w[w] = v; }
	});
	const int K = 12; const uint64_t NB = 1ull << (2 * K);
	std::vector<IDX> start(NB + 1, 0);
	auto key = [&](uint64_t p) -> uint32_t { return (uint32_t)(T.win(p) >> (64 - 2 * K)); };
	std::vector<IDX> sa(N);   // rows 1..N of the (N+1)-row matrix; row 0 is the empty suffix
	{
		// counting sort by the first K bases: one histogram per slice of positions, offsets = bucket start + what earlier slices hold
		std::vector<std::vector<IDX>> hist(nt, std::vector<IDX>(NB, 0));
		parallel_for(nt, (size_t)N, [&](int t, size_t lo, size_t hi) { std::vector<IDX>& h = hist[t]; for (size_t p = lo; p < hi; p++) h[key(p)]++; });
		IDX run = 0;
		for (uint64_t b = 0; b < NB; b++) { start[b] = run; for (int t = 0; t < nt; t++) { const IDX c = hist[t][b]; hist[t][b] = run; run += c; } }
		start[NB] = run;
		parallel_for(nt, (size_t)N, [&](int t, size_t lo, size_t hi) { std::vector<IDX>& h = hist[t]; for (size_t p = lo; p < hi; p++) sa[h[key(p)]++] = (IDX)p; });
	}
	{
		std::atomic<uint64_t> next{0}; const uint64_t chunk = 4096;
		auto work = [&]() {
			while (true)
			{
				const uint64_t b0 = next.fetch_add(chunk); if (b0 >= NB) break;
				const uint64_t b1 = std::min(NB, b0 + chunk);
				for (uint64_t b = b0; b < b1; b++)
				{
					const IDX lo = start[b], hi = start[b + 1];
					if (hi - lo < 2) continue;
					// from position 0: a suffix with fewer than K bases left shares its bucket with longer ones through the zero padding
					std::sort(sa.begin() + lo, sa.begin() + hi, [&](IDX a, IDX c) { return T.less<IDX>(a, c); });
				}
			}
		};
		std::vector<std::thread> th; for (int i = 1; i < std::max(1, threads); i++) th.emplace_back(work);
		work(); for (auto& x : th) x.join();
	}
	// ---- BWT (bwt_pac2bwt's convention: the row whose suffix is the whole text is `primary` and holds no symbol) ----
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0};
	for (uint64_t i = 0; i < L; i++) L2[fwd[i] + 1]++;
	{ const uint64_t a = L2[1], c = L2[2], g = L2[3], t = L2[4]; L2[1] = a + t; L2[2] = c + g; L2[3] = g + c; L2[4] = t + a; }   // the reverse complement adds the mirrored counts
	for (int c = 0; c < 4; c++) L2[c + 1] += L2[c];
	std::vector<uint8_t> bw(N);   // symbol of row r (r = 0..N without primary), in row order
	{
		std::vector<uint64_t> prim(nt, 0);
		parallel_for(nt, (size_t)N, [&](int t, size_t lo, size_t hi) { for (size_t r = lo; r < hi; r++) if (sa[r] == 0) prim[t] = r + 1; });
		for (int t = 0; t < nt; t++) if (prim[t]) primary = prim[t];
		bw[0] = (uint8_t)base(N - 1);   // row 0: the empty suffix, preceded by the last base
		// matrix row r + 1 holds suffix sa[r]; rows after `primary` move up by one
		parallel_for(nt, (size_t)N, [&](int, size_t lo, size_t hi) {
			for (size_t r = lo; r < hi; r++) { const IDX p = sa[r]; if (p == 0) continue; const uint64_t row = r + 1; bw[row < primary ? row : row - 1] = (uint8_t)base((uint64_t)p - 1); }
		});
	}
	fprintf(stdout, "[bwt_index] %.2f seconds elapse.\n", (float)difftime(time(NULL), w0));
	// ---- .bwt with the Occ counts interleaved every 128 symbols (bwt_bwtupdate_core) ----
	fprintf(stdout, "[bwt_index] Update BWT... "); fflush(stdout); t0 = clock();
	{
		const uint64_t n_occ = (N + 127) / 128 + 1, raw_words = (N + 15) / 16, words = raw_words + n_occ * 8;
		std::vector<uint32_t> out(words, 0);
		uint64_t c[4] = {0, 0, 0, 0}, k = 0;
		for (uint64_t i = 0; i < N; i++)
		{
			if (i % 128 == 0) { memcpy(&out[k], c, 32); k += 8; }
			if (i % 16 == 0) k++;
			out[k - 1] |= (uint32_t)bw[i] << ((15 - (i & 15)) << 1);
			c[bw[i]]++;
		}
		memcpy(&out[k], c, 32);
		FILE* fp = fopen((prefix + ".bwt").c_str(), "wb"); if (!fp) { fprintf(stdout, "\nError! cannot write %s.bwt\n", prefix.c_str()); return 1; }
		write_or_die(fp, &primary, 8); write_or_die(fp, L2 + 1, 32); write_or_die(fp, out.data(), out.size() * 4); fclose(fp);
	}
	fprintf(stdout, "%.2f sec\n", (float)(clock() - t0) / CLOCKS_PER_SEC);
	// ---- forward-only .pac, .ann, .amb (second bns_fasta2bntseq pass, for_only = 1) ----
	fprintf(stdout, "[bwt_index] Pack forward-only FASTA... "); fflush(stdout); t0 = clock();
	if (write_pac_ann_amb(prefix, fwd, anns, ambs, nullptr)) return 1;
	fprintf(stdout, "%.2f sec\n", (float)(clock() - t0) / CLOCKS_PER_SEC);
	// ---- .sa: every 32nd row of the (N+1)-row matrix (bwt_cal_sa; sa[0] is not stored) ----
	fprintf(stdout, "[bwt_index] Construct SA from BWT and Occ... "); fflush(stdout); t0 = clock();
	{
		const uint64_t intv = 32, n_sa = (N + intv) / intv;
		std::vector<uint64_t> smp(n_sa, 0);
		for (uint64_t j = 1; j < n_sa; j++) smp[j] = sa[j * intv - 1];   // row r >= 1 of the matrix is sa[r - 1]
		FILE* fp = fopen((prefix + ".sa").c_str(), "wb"); if (!fp) { fprintf(stdout, "\nError! cannot write %s.sa\n", prefix.c_str()); return 1; }
		write_or_die(fp, &primary, 8); write_or_die(fp, L2 + 1, 32); write_or_die(fp, &intv, 8); write_or_die(fp, &N, 8);
		write_or_die(fp, smp.data() + 1, (n_sa - 1) * 8); fclose(fp);
	}
	fprintf(stdout, "%.2f sec\n", (float)(clock() - t0) / CLOCKS_PER_SEC);
	return 0;
}

// .pac / .ann / .amb (second bns_fasta2bntseq pass, for_only = 1) -- shared by the host and the device builder
static int write_pac_ann_amb(const std::string& prefix, const std::vector<uint8_t>& fwd, const std::vector<Ann>& anns, const std::vector<Amb>& ambs, std::vector<uint8_t>* pac_out)
{
	const uint64_t L = fwd.size();
	std::vector<uint8_t> pac((size_t)(L >> 2) + ((L & 3) ? 1 : 0) + 1, 0);
	for (uint64_t i = 0; i < L; i++) pac[i >> 2] |= (uint8_t)(fwd[i] << ((~i & 3) << 1));
	FILE* fp = fopen((prefix + ".pac").c_str(), "wb"); if (!fp) { fprintf(stdout, "\nError! cannot write %s.pac\n", prefix.c_str()); return 1; }
	write_or_die(fp, pac.data(), (size_t)(L >> 2) + ((L & 3) ? 1 : 0));
	uint8_t ct = 0; if (L % 4 == 0) write_or_die(fp, &ct, 1);
	ct = (uint8_t)(L % 4); write_or_die(fp, &ct, 1); fclose(fp);
	fp = fopen((prefix + ".ann").c_str(), "w"); if (!fp) return 1;
	fprintf(fp, "%lld %d %u\n", (long long)L, (int)anns.size(), 11u);
	for (const Ann& a : anns)
	{
		fprintf(fp, "%d %s", 0, a.name.c_str());
		if (!a.anno.empty()) fprintf(fp, " %s\n", a.anno.c_str()); else fprintf(fp, "\n");
		fprintf(fp, "%lld %d %d\n", a.offset, a.len, a.n_ambs);
	}
	fclose(fp);
	fp = fopen((prefix + ".amb").c_str(), "w"); if (!fp) return 1;
	fprintf(fp, "%lld %d %u\n", (long long)L, (int)anns.size(), (unsigned)ambs.size());
	for (const Amb& h : ambs) fprintf(fp, "%lld %d %c\n", h.offset, h.len, h.amb);
	fclose(fp);
	if (pac_out) pac_out->swap(pac);
	return 0;
}

// .bwt and .sa from what the device built (kb_index_build): the same headers as bwt_dump_bwt / bwt_dump_sa (bwt.c:174-196)
static int write_bwt_sa(const std::string& prefix, const kb_built_index_t& b)
{
	FILE* fp = fopen((prefix + ".bwt").c_str(), "wb"); if (!fp) { fprintf(stdout, "\nError! cannot write %s.bwt\n", prefix.c_str()); return 1; }
	write_or_die(fp, &b.primary, 8); write_or_die(fp, b.L2 + 1, 32); write_or_die(fp, b.bwt, (size_t)b.bwt_words * 4); fclose(fp);
	const uint64_t intv = 32;
	fp = fopen((prefix + ".sa").c_str(), "wb"); if (!fp) { fprintf(stdout, "\nError! cannot write %s.sa\n", prefix.c_str()); return 1; }
	write_or_die(fp, &b.primary, 8); write_or_die(fp, b.L2 + 1, 32); write_or_die(fp, &intv, 8); write_or_die(fp, &b.seq_len, 8);
	write_or_die(fp, b.sa, (size_t)(b.n_sa - 1) * 8); fclose(fp);
	return 0;
}

// `kart index -gpu ref.fa prefix`: the text is packed on the host, the suffix array, BWT, Occ counts and SA samples are built on the
// device (csrc/kb_index_build.cu). `kart index -gpu -pac prefix`: the same from an existing prefix.pac / prefix.ann (l_pac is the
// first number of the .ann file), e.g. a genome that was never a FASTA file.
int build_index_gpu(const char* fa, const char* prefix_c, bool from_pac)
{
	const std::string prefix = prefix_c;
	std::vector<uint8_t> pac; long long L = 0;
	time_t w0 = time(NULL);
	if (from_pac)
	{
		FILE* fp = fopen((prefix + ".ann").c_str(), "r"); if (!fp) { fprintf(stdout, "Error! cannot open %s.ann\n", prefix.c_str()); return 1; }
		if (fscanf(fp, "%lld", &L) != 1 || L <= 0) { fclose(fp); fprintf(stdout, "Error! bad %s.ann\n", prefix.c_str()); return 1; }
		fclose(fp);
		fp = fopen((prefix + ".pac").c_str(), "rb"); if (!fp) { fprintf(stdout, "Error! cannot open %s.pac\n", prefix.c_str()); return 1; }
		pac.assign((size_t)(L / 4 + 1), 0);
		size_t got = fread(pac.data(), 1, pac.size(), fp); fclose(fp);
		if (got < (size_t)((L + 3) / 4)) { fprintf(stdout, "Error! %s.pac is too short\n", prefix.c_str()); return 1; }
		if (L % 4 == 0) pac[(size_t)(L / 4)] = 0;   // the byte behind the bases is the file's own trailer
	}
	else
	{
		fprintf(stdout, "[bwt_index] Pack FASTA... "); fflush(stdout);
		std::vector<uint8_t> fwd; std::vector<Ann> anns; std::vector<Amb> ambs;
		if (!read_reference(fa, fwd, anns, ambs)) { fprintf(stdout, "\nError! cannot open %s\n", fa); return 1; }
		if (fwd.empty()) { fprintf(stdout, "Error! %s holds no sequence\n", fa); return 1; }
		L = (long long)fwd.size();
		if (write_pac_ann_amb(prefix, fwd, anns, ambs, &pac)) return 1;
		fprintf(stdout, "%.2f sec\n", (float)difftime(time(NULL), w0));
	}
	fprintf(stdout, "[bwt_index] Construct BWT, Occ and SA on the GPU...\n"); fflush(stdout);
	kb_built_index_t b;
	int rc = kb_index_build(0, pac.data(), L, &b);
	if (rc) { fprintf(stdout, "Error! GPU index construction failed: %s (%s)\n", kb_strerror(rc), kb_index_build_error()); return 1; }
	rc = write_bwt_sa(prefix, b);
	kb_index_free(&b);
	fprintf(stdout, "[bwt_index] %.2f seconds elapse.\n", (float)difftime(time(NULL), w0));
	return rc;
}

int build_index(const char* fa, const char* prefix_c, int threads)
{
	const std::string prefix = prefix_c;
	clock_t t0 = clock();
	fprintf(stdout, "[bwt_index] Pack FASTA... "); fflush(stdout);
	std::vector<uint8_t> fwd; std::vector<Ann> anns; std::vector<Amb> ambs;
	if (!read_reference(fa, fwd, anns, ambs)) { fprintf(stdout, "\nError! cannot open %s\n", fa); return 1; }
	const uint64_t L = fwd.size(), N = 2 * L;
	fprintf(stdout, "%.2f sec\n", (float)(clock() - t0) / CLOCKS_PER_SEC);
	if (L == 0) { fprintf(stdout, "Error! %s holds no sequence\n", fa); return 1; }
	const bool wide = N + 2 >= 0xFFFFFFFFull || (getenv("KART_INDEX_64") && atoi(getenv("KART_INDEX_64")));
	return wide ? build_core<uint64_t>(prefix, fwd, anns, ambs, threads) : build_core<uint32_t>(prefix, fwd, anns, ambs, threads);
}
