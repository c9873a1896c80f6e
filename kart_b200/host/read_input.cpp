// FASTA / FASTQ (plain or .gz) chunk reader with the reference's parsing rules (src/GetData.cpp:29-49 header trimming,
// :51-107 entries, :109-143 chunks: entries are fetched two at a time, mate 2 is reverse-complemented at read time).
#include "kart_host.h"
#include <algorithm>
#include <atomic>
#include <string.h>
#include <thread>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

static size_t KB_IO_CHUNK = 32u << 20;             // bytes per gzread (KART_B200_IO_CHUNK overrides it: the tests use tiny chunks to exercise refills)

static inline char comp_base(char c)   // GetComplementaryBase, src/tools.cpp:3
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}

// ---- GzInput ---------------------------------------------------------------------------------------------------------
// BGZF block at `o`: gzip member header with FEXTRA whose extra field holds the subfield 'B','C',2,<block size - 1>. Returns the
// block's size (0: not a BGZF block) and where its deflate data starts.
static size_t bgzf_block(const uint8_t* m, size_t len, size_t o, size_t* data_off)
{
	if (o + 18 > len || m[o] != 0x1f || m[o + 1] != 0x8b || m[o + 2] != 8 || !(m[o + 3] & 4)) return 0;
	const size_t xlen = (size_t)m[o + 10] | ((size_t)m[o + 11] << 8);
	if (o + 12 + xlen > len) return 0;
	size_t bsize = 0;
	for (size_t p = o + 12; p + 4 <= o + 12 + xlen;)
	{
		const size_t slen = (size_t)m[p + 2] | ((size_t)m[p + 3] << 8);
		if (m[p] == 'B' && m[p + 1] == 'C' && slen == 2 && p + 6 <= o + 12 + xlen) bsize = ((size_t)m[p + 4] | ((size_t)m[p + 5] << 8)) + 1;
		p += 4 + slen;
	}
	if ((m[o + 3] & ~4) != 0) return 0;   // a name / comment / header CRC in front of the data: not what bgzip writes, leave it to zlib
	if (bsize < 12 + xlen + 8 || o + bsize > len) return 0;
	*data_off = o + 12 + xlen;
	return bsize;
}
bool GzInput::open(const char* path, int n_threads)
{
	close();
	threads = n_threads < 1 ? 1 : n_threads;
	int fd = ::open(path, O_RDONLY);
	if (fd >= 0)
	{
		struct stat st;
		if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size >= 28 && !getenv("KART_B200_NO_BGZF"))
		{
			void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED)
			{
				// the whole file must be one chain of BGZF blocks; anything else (plain text, ordinary gzip, trailing bytes) is zlib's
				const uint8_t* mm = (const uint8_t*)m; size_t o = 0, d;
				while (o < (size_t)st.st_size) { const size_t b = bgzf_block(mm, (size_t)st.st_size, o, &d); if (!b) break; o += b; }
				if (o == (size_t)st.st_size) { map = mm; map_len = o; at = 0; bgzf = true; madvise(m, map_len, MADV_SEQUENTIAL); }
				else munmap(m, (size_t)st.st_size);
			}
		}
		::close(fd);
	}
	if (bgzf) return true;
	fp = gzopen(path, "rb");
	if (!fp) return false;
	gzbuffer(fp, 1 << 20);
	return true;
}
void GzInput::close()
{
	if (fp) gzclose(fp);
	if (map) munmap((void*)map, map_len);
	fp = nullptr; map = nullptr; map_len = at = 0; carry.clear(); carry_pos = 0; bgzf = false; failed = false;
}
static bool bgzf_inflate(const uint8_t* m, size_t data_off, size_t block_end, char* dst, size_t isize)
{
	const uint8_t* tail = m + block_end - 8;
	const uint32_t crc = (uint32_t)tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
	z_stream zs; memset(&zs, 0, sizeof(zs));
	if (inflateInit2(&zs, -15) != Z_OK) return false;
	zs.next_in = (Bytef*)(m + data_off); zs.avail_in = (uInt)(block_end - 8 - data_off);
	zs.next_out = (Bytef*)dst; zs.avail_out = (uInt)isize;
	const int rc = inflate(&zs, Z_FINISH);
	const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
	inflateEnd(&zs);
	return ok && (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)dst, (uInt)isize) == crc;
}
int GzInput::read(char* dst, unsigned n)
{
	if (!bgzf) return fp ? gzread(fp, dst, n) : -1;
	if (failed) return -1;
	if (carry_pos < carry.size())
	{
		const size_t k = std::min((size_t)n, carry.size() - carry_pos);
		memcpy(dst, carry.data() + carry_pos, k); carry_pos += k;
		return (int)k;
	}
	// as many whole blocks as fit into n bytes
	struct Blk { size_t data, end, out, isize; };
	std::vector<Blk> blk; size_t total = 0;
	while (at < map_len && blk.size() < 65536)
	{
		size_t d; const size_t b = bgzf_block(map, map_len, at, &d);
		const uint8_t* t = map + at + b - 4;
		const size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
		if (isize == 0) { at += b; continue; }                 // empty blocks (the end-of-file marker) hold nothing
		if (total + isize > (size_t)n) break;
		blk.push_back({d, at + b, total, isize}); total += isize; at += b;
	}
	if (blk.empty())
	{
		if (at >= map_len) return 0;
		// the next block is larger than the request: through the carry buffer
		size_t d; const size_t b = bgzf_block(map, map_len, at, &d);
		const uint8_t* t = map + at + b - 4;
		const size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
		carry.resize(isize); carry_pos = 0;
		if (!bgzf_inflate(map, d, at + b, carry.data(), isize)) { carry.clear(); failed = true; return -1; }
		at += b;
		return read(dst, n);
	}
	std::atomic<size_t> first_bad{blk.size()};
	parallel_for(threads, blk.size(), [&](int, size_t lo, size_t hi) {
		for (size_t i = lo; i < hi; i++)
			if (!bgzf_inflate(map, blk[i].data, blk[i].end, dst + blk[i].out, blk[i].isize)) { size_t cur = first_bad.load(); while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {} }
	});
	if (first_bad.load() == blk.size()) return (int)total;
	// a damaged block: what lies in front of it is delivered (gzread does the same), then the input has ended in an error
	at = map_len; failed = true;
	const size_t good = blk[first_bad.load()].out;
	return good ? (int)good : -1;
}

bool ReadSource::open(const char* f1, const char* f2)
{
	{ gzFile t = gzopen(f1, "rb"); if (!t) return false; char c = 0; gzread(t, &c, 1); gzclose(t); fastq = (c == '@'); }   // CheckReadFormat :8
	for (Stream* s : {&s1, &s2}) { s->in.close(); s->pos = s->end = 0; s->eof = false; s->pending.clear(); s->has_pending = false; s->lo = s->hi = 0; s->nl.clear(); s->nl_used = 0; s->drained = false; }
	two = f2 != nullptr;
	{ const char* e = getenv("KART_B200_IO_CHUNK"); if (e && atol(e) >= 16) KB_IO_CHUNK = (size_t)atol(e); }
	const int per = threads > 1 && two ? (threads + 1) / 2 : threads;
	if (!s1.in.open(f1, per)) return false;
	s1.buf.resize(1 << 22);
	if (two) { if (!s2.in.open(f2, per)) { s1.in.close(); return false; } s2.buf.resize(1 << 22); }
	return true;
}

void ReadSource::close() { s1.in.close(); s2.in.close(); }

bool ReadSource::line(Stream& s, std::string& out)   // one line including its '\n' (getline semantics)
{
	if (s.has_pending) { out.swap(s.pending); s.has_pending = false; return true; }
	out.clear();
	while (true)
	{
		if (s.pos == s.end)
		{
			if (s.eof) return !out.empty();
			int got = s.in.read(s.buf.data(), (unsigned)s.buf.size());
			if (got <= 0) { s.eof = true; return !out.empty(); }
			s.pos = 0; s.end = (size_t)got;
		}
		const char* p = s.buf.data() + s.pos; size_t avail = s.end - s.pos;
		const char* nl = (const char*)memchr(p, '\n', avail);
		if (nl) { out.append(p, nl - p + 1); s.pos += nl - p + 1; return true; }
		out.append(p, avail); s.pos = s.end;
	}
}

bool ReadSource::entry(Stream& s, std::string& name, std::string& seq, std::string& qual)   // GetNextEntry :51 ; false <=> rlen == 0
{
	std::string ln; name.clear(); seq.clear(); qual.clear();
	if (!line(s, ln)) return false;
	int len = (int)ln.size(), p1 = len - 1, p2 = len - 1;
	for (int i = 1; i < len; i++) if (ln[i] != '>' && ln[i] != '@') { p1 = i; break; }                 // IdentifyHeaderBegPos :29
	for (int i = 1; i < len; i++) if (ln[i] == ' ' || ln[i] == '/' || ln[i] == '\t') { p2 = i; break; }   // IdentifyHeaderEndPos :40
	if (p2 > p1) name.assign(ln, p1, p2 - p1);
	if (fastq)
	{
		std::string plus;
		if (!line(s, seq)) return false;
		line(s, plus); line(s, qual);
		int rl = (int)seq.size() - 1;                       // rlen = getline length - 1 (:70-76)
		if (rl <= 0) { seq.clear(); return false; }
		seq.resize(rl);
		qual.resize(rl + 1, '\0'); qual.resize(rl);
		size_t z = qual.find('\0'); if (z != std::string::npos) qual.resize(z);
	}
	else
	{
		while (line(s, ln))
		{
			if (ln[0] == '>') { s.pending.swap(ln); s.has_pending = true; break; }
			seq.append(ln, 0, ln.size() - 1);
		}
		if (seq.empty()) return false;
	}
	return true;
}

int ReadSource::fill_serial(ReadBatch& b, int max_reads, bool pair_end)
{
	int added = 0; std::string name, seq, qual;
	auto push = [&](bool flip) {
		if (flip)
		{
			std::string rc(seq.size(), 'N');
			for (size_t i = 0; i < seq.size(); i++) rc[seq.size() - 1 - i] = comp_base(seq[i]);
			seq.swap(rc);
			if (fastq) std::reverse(qual.begin(), qual.end());
		}
		b.seq.append((const uint8_t*)seq.data(), seq.size()); b.seq_off.push_back(b.seq.size());
		if (fastq) { qual.resize(seq.size(), ' '); b.qual.append(qual.data(), qual.size()); }
		b.names.append(name.data(), name.size()); b.name_off.push_back((uint32_t)b.names.size());
		added++;
	};
	while (added + 2 <= max_reads || added == 0)
	{
		if (!entry(s1, name, seq, qual)) break;
		push(false);
		if (!entry(two ? s2 : s1, name, seq, qual)) break;
		push(pair_end);
		if (added >= max_reads) break;
	}
	return added;
}

// ------------------------------------------------------------------------------------------------
// Block path (FASTQ): same records as fill_serial, produced by `threads` workers.
//
// A FASTQ entry of the reference parser is exactly four getline() calls (GetData.cpp:61-76), so entry k of a stream starts at
// line 4k: once the newline offsets of a buffered piece of text are known, entries are independent. Every stream keeps
// text[lo,hi) and the offsets nl[] of its newlines; workers index new text in slices, then two passes over the entries:
// pass A measures (read length, name extent) per entry, pass B copies sequence / quality / name to their final offsets
// (mate 2 reverse-complemented, quality reversed). Whatever is not a complete, well-formed four-line entry -- the tail of the
// input without a final newline, an entry with an empty sequence line -- goes through the entry-at-a-time code below,
// which follows fill_serial statement by statement.
// ------------------------------------------------------------------------------------------------
void parallel_for(int threads, size_t n, const std::function<void(int, size_t, size_t)>& body)
{
	if (threads < 1) threads = 1;
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	if (threads == 1) { body(0, 0, n); return; }
	std::vector<std::thread> th;
	for (int t = 1; t < threads; t++) th.emplace_back([&, t]() { body(t, n * t / threads, n * (t + 1) / threads); });
	body(0, 0, n / threads);
	for (auto& x : th) x.join();
}

static const size_t KB_IO_MAX_TEXT = 1u << 30;     // soft cap of buffered text per stream (offsets are 32 bit)

static void index_newlines(ReadSource::Stream& s, size_t from, size_t to, int threads)
{
	if (threads < 1) threads = 1;
	std::vector<std::vector<uint32_t>> found(threads);
	const char* t = s.text.data();
	parallel_for(threads, to - from, [&](int w, size_t a, size_t b) {
		std::vector<uint32_t>& v = found[w]; v.reserve((b - a) / 48 + 16);
		const char* p = t + from + a; const char* e = t + from + b;
		while (p < e) { const char* q = (const char*)memchr(p, '\n', e - p); if (!q) break; v.push_back((uint32_t)(q - t)); p = q + 1; }
	});
	size_t add = 0; for (auto& v : found) add += v.size();
	size_t at = s.nl.size(); s.nl.resize(at + add);
	for (auto& v : found) { if (!v.empty()) memcpy(s.nl.data() + at, v.data(), v.size() * 4); at += v.size(); }
}

// Makes `want` complete entries (4 lines each) available after the cursor, or everything there is when the input ends first.
size_t ReadSource::buffer_records(Stream& s, size_t want)
{
	while (true)
	{
		size_t avail = (s.nl.size() - s.nl_used) / 4;
		if (avail >= want || s.drained) return avail < want ? avail : want;
		if (avail >= 1 && s.hi - s.lo >= KB_IO_MAX_TEXT) return avail;
		if (s.lo > 0)
		{
			// drop the consumed text, keep the newline offsets of what remains
			size_t keep = s.hi - s.lo, nk = s.nl.size() - s.nl_used;
			memmove(s.text.data(), s.text.data() + s.lo, keep);
			uint32_t* nl = s.nl.data();
			for (size_t i = 0; i < nk; i++) nl[i] = nl[s.nl_used + i] - (uint32_t)s.lo;
			s.nl.resize(nk); s.nl_used = 0; s.hi = keep; s.lo = 0;
		}
		if (s.hi + KB_IO_CHUNK > 0xFFFFFF00u) { s.drained = true; continue; }      // an "entry" beyond 4 GB: give up on the stream like a read error
		s.text.n = s.hi; if (s.text.cap < s.hi + KB_IO_CHUNK) s.text.reserve(std::max(s.hi + KB_IO_CHUNK, std::min((size_t)KB_IO_MAX_TEXT, want * 512) + 2 * KB_IO_CHUNK));
		int got = s.in.read(s.text.data() + s.hi, (unsigned)KB_IO_CHUNK);
		if (got <= 0) { s.drained = true; continue; }
		index_newlines(s, s.hi, s.hi + (size_t)got, threads > 1 && two ? (threads + 1) / 2 : threads);
		s.hi += (size_t)got;
	}
}

static uint8_t g_comp[256]; static bool g_comp_init = false;
static void comp_init() { if (!g_comp_init) { for (int c = 0; c < 256; c++) g_comp[c] = (uint8_t)comp_base((char)c); g_comp_init = true; } }

// start offset of line m of a stream's index (m counts from the first indexed line)
static inline size_t line_start(const ReadSource::Stream& s, size_t m, size_t first) { return m == 0 ? first : (size_t)s.nl[m - 1] + 1; }

// next line after the cursor, getline semantics on the buffered text; false at the end of the input
static bool mem_line(ReadSource::Stream& s, const char*& p, size_t& len, const std::function<size_t(ReadSource::Stream&, size_t)>& more)
{
	if (s.nl_used == s.nl.size() && !s.drained) more(s, 1);
	if (s.nl_used < s.nl.size()) { size_t e = (size_t)s.nl[s.nl_used] + 1; p = s.text.data() + s.lo; len = e - s.lo; s.lo = e; s.nl_used++; return true; }
	if (s.lo < s.hi) { p = s.text.data() + s.lo; len = s.hi - s.lo; s.lo = s.hi; return true; }
	return false;
}

int ReadSource::fill_blocks(ReadBatch& b, int max_reads, bool pair_end)
{
	comp_init();
	int added = 0;
	Stream& sa = s1; Stream& sb = two ? s2 : s1;
	std::vector<uint32_t> rl, nb, nlen;    // per entry: read length, name offset in its stream's text, name length
	while (added + 2 <= max_reads)
	{
		size_t want = (size_t)(max_reads - added) / 2, P;
		if (two)
		{
			size_t n1 = 0, n2 = 0;
			std::thread t2([&]() { n2 = buffer_records(s2, want); });
			n1 = buffer_records(s1, want); t2.join();
			P = n1 < n2 ? n1 : n2;
		}
		else P = buffer_records(s1, want * 2) / 2;
		if (P == 0) break;
		const size_t E = 2 * P;
		rl.resize(E); nb.resize(E); nlen.resize(E);
		// where line 0 of the index starts: the cursor when nothing of the index is consumed yet
		const size_t first_a = sa.nl_used == 0 ? sa.lo : 0, first_b = sb.nl_used == 0 ? sb.lo : 0;
		const size_t base_a = sa.nl_used, base_b = sb.nl_used;
		auto entry_line0 = [&](size_t e, const Stream*& s, size_t& first) -> size_t {
			if (two) { if (e & 1) { s = &sb; first = first_b; return base_b + 4 * (e >> 1); } s = &sa; first = first_a; return base_a + 4 * (e >> 1); }
			s = &sa; first = first_a; return base_a + 4 * e;
		};
		int W = threads < 1 ? 1 : threads;
		std::vector<uint64_t> sum_seq(W + 1, 0), sum_name(W + 1, 0); std::vector<size_t> bad(W, (size_t)-1);
		// pass A
		parallel_for(W, P, [&](int w, size_t lo, size_t hi) {
			uint64_t ss = 0, sn = 0;
			for (size_t e = 2 * lo; e < 2 * hi; e++)
			{
				const Stream* s; size_t first; size_t m = entry_line0(e, s, first);
				size_t h0 = line_start(*s, m, first), h1 = (size_t)s->nl[m] + 1, q1 = (size_t)s->nl[m + 1] + 1;
				const char* ln = s->text.data() + h0; int len = (int)(h1 - h0), p1 = len - 1, p2 = len - 1;
				for (int i = 1; i < len; i++) if (ln[i] != '>' && ln[i] != '@') { p1 = i; break; }
				for (int i = 1; i < len; i++) if (ln[i] == ' ' || ln[i] == '/' || ln[i] == '\t') { p2 = i; break; }
				uint32_t r = (uint32_t)(q1 - h1 - 1);
				if (r == 0) { if (bad[w] == (size_t)-1) bad[w] = e; }
				rl[e] = r; nb[e] = (uint32_t)(h0 + p1); nlen[e] = p2 > p1 ? (uint32_t)(p2 - p1) : 0;
				ss += r; sn += nlen[e];
			}
			sum_seq[w + 1] = ss; sum_name[w + 1] = sn;
		});
		size_t fail = (size_t)-1; for (int w = 0; w < W; w++) if (bad[w] < fail) fail = bad[w];
		if (fail != (size_t)-1)
		{
			// an entry with an empty sequence line ends the chunk (GetNextChunk: rlen == 0): keep the pairs before it here, the rest is the serial code's business
			P = fail / 2;
			std::fill(sum_seq.begin(), sum_seq.end(), 0); std::fill(sum_name.begin(), sum_name.end(), 0);
			parallel_for(W, P, [&](int w, size_t lo, size_t hi) { uint64_t ss = 0, sn = 0; for (size_t e = 2 * lo; e < 2 * hi; e++) { ss += rl[e]; sn += nlen[e]; } sum_seq[w + 1] = ss; sum_name[w + 1] = sn; });
		}
		for (int w = 0; w < W; w++) { sum_seq[w + 1] += sum_seq[w]; sum_name[w + 1] += sum_name[w]; }
		if (P > 0)
		{
			const size_t seq0 = b.seq.size(), name0 = b.names.size(), r0 = (size_t)b.n();
			b.seq.resize(seq0 + sum_seq[W]); b.qual.resize(seq0 + sum_seq[W]); b.names.resize(name0 + sum_name[W]);
			b.seq_off.resize(r0 + 1 + 2 * P); b.name_off.resize(r0 + 1 + 2 * P);
			// pass B (the slices are the same as in pass A: parallel_for cuts [0,P) deterministically)
			parallel_for(W, P, [&](int w, size_t lo, size_t hi) {
				size_t so = seq0 + sum_seq[w], no = name0 + sum_name[w];
				uint8_t* seq = b.seq.data(); char* qual = b.qual.data(); char* names = b.names.data();
				for (size_t e = 2 * lo; e < 2 * hi; e++)
				{
					const Stream* s; size_t first; size_t m = entry_line0(e, s, first);
					const char* t = s->text.data();
					const uint8_t* sq = (const uint8_t*)t + (size_t)s->nl[m] + 1; size_t r = rl[e];
					const char* ql = t + (size_t)s->nl[m + 2] + 1; size_t l3 = (size_t)s->nl[m + 3] + 1 - ((size_t)s->nl[m + 2] + 1);
					size_t q = r < l3 ? r : l3; const void* z = memchr(ql, 0, q); if (z) q = (size_t)((const char*)z - ql);
					bool flip = pair_end && (e & 1);
					if (!flip) { memcpy(seq + so, sq, r); memcpy(qual + so, ql, q); }
					else { for (size_t i = 0; i < r; i++) seq[so + i] = g_comp[sq[r - 1 - i]]; for (size_t i = 0; i < q; i++) qual[so + i] = ql[q - 1 - i]; }
					if (q < r) memset(qual + so + q, ' ', r - q);
					memcpy(names + no, t + nb[e], nlen[e]);
					so += r; no += nlen[e];
					b.seq_off[r0 + 1 + e] = so; b.name_off[r0 + 1 + e] = (uint32_t)no;
				}
			});
			// advance the cursors past the consumed entries
			if (two) { sa.nl_used += 4 * P; sa.lo = (size_t)sa.nl[sa.nl_used - 1] + 1; sb.nl_used += 4 * P; sb.lo = (size_t)sb.nl[sb.nl_used - 1] + 1; }
			else { sa.nl_used += 8 * P; sa.lo = (size_t)sa.nl[sa.nl_used - 1] + 1; }
			added += (int)(2 * P);
		}
		if (fail != (size_t)-1) break;
	}
	// entry-at-a-time: the end of the input (last line without '\n', odd entry counts) and malformed entries. Same statements as fill_serial.
	std::function<size_t(Stream&, size_t)> more = [this](Stream& s, size_t w) { return buffer_records(s, w); };
	auto entry_mem = [&](Stream& s, std::string& name, std::string& seq, std::string& qual) -> bool {
		const char* p; size_t len; name.clear(); seq.clear(); qual.clear();
		if (!mem_line(s, p, len, more)) return false;
		int n = (int)len, p1 = n - 1, p2 = n - 1;
		for (int i = 1; i < n; i++) if (p[i] != '>' && p[i] != '@') { p1 = i; break; }
		for (int i = 1; i < n; i++) if (p[i] == ' ' || p[i] == '/' || p[i] == '\t') { p2 = i; break; }
		if (p2 > p1) name.assign(p + p1, p2 - p1);
		if (!mem_line(s, p, len, more)) return false;
		seq.assign(p, len);
		if (mem_line(s, p, len, more) && mem_line(s, p, len, more)) qual.assign(p, len);
		int r = (int)seq.size() - 1;
		if (r <= 0) { seq.clear(); return false; }
		seq.resize(r);
		qual.resize(r + 1, '\0'); qual.resize(r);
		size_t z = qual.find('\0'); if (z != std::string::npos) qual.resize(z);
		return true;
	};
	std::string name, seq, qual;
	auto push = [&](bool flip) {
		if (flip)
		{
			std::string rc(seq.size(), 'N');
			for (size_t i = 0; i < seq.size(); i++) rc[seq.size() - 1 - i] = comp_base(seq[i]);
			seq.swap(rc);
			std::reverse(qual.begin(), qual.end());
		}
		b.seq.append((const uint8_t*)seq.data(), seq.size()); b.seq_off.push_back(b.seq.size());
		qual.resize(seq.size(), ' '); b.qual.append(qual.data(), qual.size());
		b.names.append(name.data(), name.size()); b.name_off.push_back((uint32_t)b.names.size());
		added++;
	};
	while (added + 2 <= max_reads || added == 0)
	{
		if (!entry_mem(sa, name, seq, qual)) break;
		push(false);
		if (!entry_mem(sb, name, seq, qual)) break;
		push(pair_end);
		if (added >= max_reads) break;
	}
	return added;
}

int ReadSource::fill(ReadBatch& b, int max_reads, bool pair_end)
{
	b.fastq = fastq; b.pair_end = pair_end;
	return fastq ? fill_blocks(b, max_reads, pair_end) : fill_serial(b, max_reads, pair_end);
}

static std::atomic<bool> g_cuda_ready{false};
void host_cuda_ready() { g_cuda_ready.store(true); }
void* host_buf_alloc(size_t bytes, bool want_pinned, bool* got_pinned)
{
	void* p = (want_pinned && g_cuda_ready.load()) ? kb_host_alloc(bytes ? bytes : 1) : nullptr;   // cudaHostAlloc would block on the context being created
	*got_pinned = p != nullptr;
	if (!p && bytes >= (4u << 20))
	{
		// large pageable buffers: 2 MB alignment + MADV_HUGEPAGE (first-touch faults of 4 KB pages cost more than the parsing itself)
		if (posix_memalign(&p, 2u << 20, (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1)) != 0) p = nullptr;
		else madvise(p, (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1), MADV_HUGEPAGE);
	}
	if (!p) p = malloc(bytes ? bytes : 1);
	if (!p) { fprintf(stderr, "Error! out of host memory (%zu bytes)\n", bytes); exit(1); }
	return p;
}
void host_buf_free(void* p, bool pinned) { if (pinned) kb_host_free(p); else free(p); }
