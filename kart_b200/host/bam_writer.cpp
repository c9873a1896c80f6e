// BAM output (-bo), byte-compatible with what the reference produces through its vendored htslib 1.5
// (src/Mapping.cpp:610-621 sam_parse1 + sam_write1 per SAM line, :656,:676-680 header, :734 close).
//
// Freshly written: records are encoded straight from kb_aln_t (no SAM text round trip) into the byte stream that
// bam_write1 (htslib sam.c:427) would hand to bgzf_write, and the BGZF layer reproduces htslib's block policy
// (bgzf.c:1370 bgzf_flush_try before every record, :1376 bgzf_write cutting at 0xff00 bytes, :346 one raw deflate
// stream per block at the default level, :1434 empty EOF block). A block is a contiguous range of that byte stream, so
// the cut positions are found serially (cheap) and the blocks are deflated by a pool of threads and written in order --
// same bytes as the single-threaded reference given the same zlib.
#include "kart_host.h"
#include <string.h>
#include <thread>
#include <atomic>

static const size_t BGZF_BLOCK = 0xff00;

static inline void put_u32(std::string& o, uint32_t v) { char b[4] = {(char)(v & 255), (char)((v >> 8) & 255), (char)((v >> 16) & 255), (char)(v >> 24)}; o.append(b, 4); }

// hts_reg2bin(beg, end, 14, 5) (htslib/hts.h), signed arithmetic shifts like the original
static int reg2bin(int64_t beg, int64_t end)
{
	int l, s = 14, t = ((1 << 15) - 1) / 7;
	for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << (l * 3)) if ((beg >> s) == (end >> s)) return t + (int)(beg >> s);
	return 0;
}

// seq_nt16_table (htslib hts.c): "=ACMGRSVTWYHKDBN" case-insensitively, '0'..'3' as ACGT, everything else 15
static uint8_t nt16(unsigned char c)
{
	static uint8_t tab[256]; static bool init = false;
	if (!init)
	{
		memset(tab, 15, sizeof(tab));
		const char* s = "=ACMGRSVTWYHKDBN";
		for (int i = 0; i < 16; i++) { tab[(unsigned char)s[i]] = (uint8_t)i; if (s[i] >= 'A' && s[i] <= 'Z') tab[(unsigned char)(s[i] + 32)] = (uint8_t)i; }
		tab['0'] = 1; tab['1'] = 2; tab['2'] = 4; tab['3'] = 8;
		init = true;
	}
	return tab[c];
}
void bam_tables_init() { nt16('A'); }

static inline char comp_base(char c)
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}

// TAG:i:value with the integer type sam_parse1 picks (sam.c:1075-1103)
static void put_aux_int(std::string& o, const char* tag, long long v)
{
	o.append(tag, 2);
	if (v < 0)
	{
		if (v >= -128) { o += 'c'; o += (char)v; }
		else if (v >= -32768) { o += 's'; o += (char)(v & 255); o += (char)((v >> 8) & 255); }
		else { o += 'i'; put_u32(o, (uint32_t)(int32_t)v); }
	}
	else
	{
		if (v <= 255) { o += 'C'; o += (char)v; }
		else if (v <= 65535) { o += 'S'; o += (char)(v & 255); o += (char)((v >> 8) & 255); }
		else { o += 'I'; put_u32(o, (uint32_t)v); }
	}
}

// One record = what sam_parse1 makes of the SAM line sam_read_line() prints for the same arguments, serialised like bam_write1.
// Lines sam_parse1 rejects (query name longer than 252, SEQ/QUAL length mismatch) are dropped, as in the reference (:618).
void bam_read_record(std::string& o, std::vector<uint32_t>& rec_end, const std::vector<int32_t>& name2id, const ReadBatch& b, int r, bool stored_fwd, const kb_aln_t& a, const uint32_t* cigar, bool fastq)
{
	if (a.kind == 2) return;
	const char* name = b.names.data() + b.name_off[r]; size_t nlen = b.name_off[r + 1] - b.name_off[r];
	const char* seq = (const char*)b.seq.data() + b.seq_off[r]; size_t rlen = (size_t)(b.seq_off[r + 1] - b.seq_off[r]);
	const char* qual = fastq ? b.qual.data() + b.seq_off[r] : nullptr;
	{ size_t z = strnlen(name, nlen); nlen = z; }                     // the SAM line is a C string
	if (nlen + 1 > 252) return;                                       // "query name too long" (p - q counts the separator)
	size_t qlen = fastq ? strnlen(qual, rlen) : 0;
	if (fastq && !(qlen == 1 && qual[0] == '*') && qlen != rlen) return;   // "SEQ and QUAL are of different length"
	int32_t tid = -1, pos = -1, mtid = -1, mpos = -1, isize = 0; uint32_t flag = (uint32_t)a.flag, mapq = 0, n_cigar = 0;
	int64_t span = 1;
	const uint32_t* cg = nullptr;
	if (a.kind == 1)
	{
		tid = name2id[a.chr]; pos = (int32_t)(a.pos - 1); mapq = (uint32_t)a.mapq;
		if (pos < 0 && tid >= 0) tid = -1;
		if (tid < 0) flag |= 4;
		n_cigar = (uint32_t)a.cig_len; cg = cigar + a.cig_off;
		if (n_cigar >= 65536) return;                                   // bam_write1 refuses (sam.c:432)
		if (!(flag & 4)) { span = 0; for (uint32_t k = 0; k < n_cigar; k++) { uint32_t op = cg[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += cg[k] >> 4; } }
		if (n_cigar == 0) flag |= 4;
		else { size_t q = 0; for (uint32_t k = 0; k < n_cigar; k++) { uint32_t op = cg[k] & 15; if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q += cg[k] >> 4; } if (q != rlen) return; }   // "CIGAR and query sequence are of different length"
		if (a.mate_pos >= 0) { mtid = tid; mpos = (int32_t)(a.mate_pos - 1); isize = a.tlen; if (mpos < 0 && mtid >= 0) mtid = -1; }
	}
	else flag |= 4;
	uint32_t bin = (uint32_t)reg2bin(pos, pos + span) & 0xFFFF;
	size_t at = o.size();
	put_u32(o, 0);                                                     // block_len, patched below
	put_u32(o, (uint32_t)tid); put_u32(o, (uint32_t)pos);
	put_u32(o, bin << 16 | (mapq & 255) << 8 | (uint32_t)(nlen + 1));
	put_u32(o, (flag & 0xFFFF) << 16 | n_cigar);
	put_u32(o, (uint32_t)rlen); put_u32(o, (uint32_t)mtid); put_u32(o, (uint32_t)mpos); put_u32(o, (uint32_t)isize);
	o.append(name, nlen); o += '\0';
	for (uint32_t k = 0; k < n_cigar; k++) put_u32(o, cg[k]);
	bool as_is = a.kind == 0 || ((a.fwd != 0) == stored_fwd);
	size_t sq = o.size(); o.resize(sq + (rlen + 1) / 2, '\0');
	if (as_is) for (size_t i = 0; i < rlen; i++) o[sq + (i >> 1)] |= (char)(nt16((unsigned char)seq[i]) << ((~i & 1) << 2));
	else for (size_t i = 0; i < rlen; i++) o[sq + (i >> 1)] |= (char)(nt16((unsigned char)comp_base(seq[rlen - 1 - i])) << ((~i & 1) << 2));
	size_t ql = o.size(); o.resize(ql + rlen);
	if (!fastq || (qlen == 1 && qual[0] == '*')) memset(&o[ql], 0xff, rlen);
	else if (as_is) for (size_t i = 0; i < rlen; i++) o[ql + i] = (char)(qual[i] - 33);
	else for (size_t i = 0; i < rlen; i++) o[ql + i] = (char)(qual[rlen - 1 - i] - 33);
	if (a.kind == 1) put_aux_int(o, "NM", (long long)rlen - a.score);
	put_aux_int(o, "AS", a.kind == 1 ? a.score : 0); put_aux_int(o, "XS", a.kind == 1 ? a.sub_score : 0);
	uint32_t block_len = (uint32_t)(o.size() - at - 4);
	o[at] = (char)(block_len & 255); o[at + 1] = (char)((block_len >> 8) & 255); o[at + 2] = (char)((block_len >> 16) & 255); o[at + 3] = (char)(block_len >> 24);
	rec_end.push_back((uint32_t)o.size());
}

// bgzf_compress (bgzf.c:346): 18-byte header, raw deflate of the whole block in one call, CRC32 + ISIZE
static void bgzf_block(std::string& out, const char* src, size_t len, int level)
{
	static const unsigned char magic[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
	out.assign(0x10000, '\0');
	z_stream zs; memset(&zs, 0, sizeof(zs));
	zs.next_in = (Bytef*)src; zs.avail_in = (uInt)len; zs.next_out = (Bytef*)&out[18]; zs.avail_out = 0x10000 - 18 - 8;
	deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
	deflate(&zs, Z_FINISH);
	deflateEnd(&zs);
	size_t dlen = zs.total_out + 26;
	memcpy(&out[0], magic, 18);
	out[16] = (char)((dlen - 1) & 255); out[17] = (char)(((dlen - 1) >> 8) & 255);
	uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0L), (const Bytef*)src, (uInt)len);
	for (int i = 0; i < 4; i++) { out[dlen - 8 + i] = (char)((crc >> (8 * i)) & 255); out[dlen - 4 + i] = (char)((len >> (8 * i)) & 255); }
	out.resize(dlen);
}

bool BamWriter::open(const char* path, int n_threads)
{
	fp = fopen(path, "wb"); threads = n_threads > 0 ? n_threads : 1; pend.clear();
	return fp != nullptr;
}

// deflates stream[cuts[k], cuts[k+1]) for all k in parallel and writes the blocks in order
void BamWriter::emit(const std::string& stream, const std::vector<size_t>& cuts)
{
	size_t nb = cuts.size() > 0 ? cuts.size() - 1 : 0;
	if (nb == 0) return;
	std::vector<std::string> outb(nb); std::atomic<size_t> next(0);
	auto work = [&]() { for (size_t k; (k = next.fetch_add(1)) < nb;) bgzf_block(outb[k], stream.data() + cuts[k], cuts[k + 1] - cuts[k], Z_DEFAULT_COMPRESSION); };
	int nt = (int)(nb < (size_t)threads ? nb : (size_t)threads);
	if (nt <= 1) work(); else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work); for (auto& t : th) t.join(); }
	for (auto& s : outb) fwrite(s.data(), 1, s.size(), fp);
}

// Appends `bytes` written as consecutive bgzf_write units ending at unit_end[] (each preceded by bgzf_flush_try when
// `records`); `pend` holds the bytes of the block that is still open.
void BamWriter::append(const std::string& bytes, const std::vector<uint32_t>& unit_end, bool records)
{
	std::string stream; stream.reserve(pend.size() + bytes.size());
	stream = pend; size_t base = stream.size(); stream += bytes;
	std::vector<size_t> cuts(1, 0); size_t start = 0;   // start of the open block in `stream`
	size_t prev = 0;
	for (size_t u = 0; u < unit_end.size(); u++)
	{
		size_t ub = base + prev, ue = base + unit_end[u], size = ue - ub; prev = unit_end[u];
		size_t off = ub - start;
		if (records && off + size > BGZF_BLOCK && off > 0) { cuts.push_back(ub); start = ub; off = 0; }
		// bgzf_write: fill the block, flush whenever it reaches exactly 0xff00 bytes
		size_t p = ub;
		while (p < ue)
		{
			size_t c = BGZF_BLOCK - (p - start); if (c > ue - p) c = ue - p;
			p += c;
			if (p - start == BGZF_BLOCK) { cuts.push_back(p); start = p; }
		}
	}
	if (cuts.back() != start) cuts.push_back(start);
	emit(stream, cuts);
	pend.assign(stream, start, std::string::npos);
}

void BamWriter::flush()
{
	if (pend.empty()) return;
	std::vector<size_t> cuts = {0, pend.size()};
	emit(pend, cuts); pend.clear();
}

// bam_hdr_write (sam.c:228): magic, l_text, text, n_targets, per target l_name, name, l_ref ; then bgzf_flush
void BamWriter::write_header(const std::string& text, const std::vector<std::string>& names, const std::vector<int64_t>& lens)
{
	std::string o; std::vector<uint32_t> ends;
	auto unit = [&]() { ends.push_back((uint32_t)o.size()); };
	o.append("BAM\1", 4); unit();
	put_u32(o, (uint32_t)text.size()); unit();
	if (!text.empty()) { o += text; unit(); }
	put_u32(o, (uint32_t)names.size()); unit();
	for (size_t i = 0; i < names.size(); i++)
	{
		put_u32(o, (uint32_t)names[i].size() + 1); unit();
		o.append(names[i].c_str(), names[i].size() + 1); unit();
		put_u32(o, (uint32_t)lens[i]); unit();
	}
	append(o, ends, false);
	flush();
}

bool BamWriter::close()
{
	if (!fp) return false;
	flush();
	std::string eof; bgzf_block(eof, "", 0, Z_DEFAULT_COMPRESSION);
	fwrite(eof.data(), 1, eof.size(), fp);
	bool ok = fclose(fp) == 0; fp = nullptr;
	return ok;
}
