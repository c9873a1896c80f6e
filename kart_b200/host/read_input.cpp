// FASTA / FASTQ (plain or .gz) chunk reader with the reference's parsing rules (src/GetData.cpp:29-49 header trimming,
// :51-107 entries, :109-143 chunks: entries are fetched two at a time, mate 2 is reverse-complemented at read time).
#include "kart_host.h"
#include <algorithm>
#include <string.h>

static inline char comp_base(char c)   // GetComplementaryBase, src/tools.cpp:3
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}

bool ReadSource::open(const char* f1, const char* f2)
{
	{ gzFile t = gzopen(f1, "rb"); if (!t) return false; char c = 0; gzread(t, &c, 1); gzclose(t); fastq = (c == '@'); }   // CheckReadFormat :8
	s1 = Stream(); s2 = Stream(); two = f2 != nullptr;
	s1.fp = gzopen(f1, "rb"); if (!s1.fp) return false; gzbuffer(s1.fp, 1 << 20); s1.buf.resize(1 << 22);
	if (two) { s2.fp = gzopen(f2, "rb"); if (!s2.fp) { gzclose(s1.fp); s1.fp = nullptr; return false; } gzbuffer(s2.fp, 1 << 20); s2.buf.resize(1 << 22); }
	return true;
}

void ReadSource::close() { if (s1.fp) gzclose(s1.fp); if (s2.fp) gzclose(s2.fp); s1.fp = s2.fp = nullptr; }

bool ReadSource::line(Stream& s, std::string& out)   // one line including its '\n' (getline semantics)
{
	if (s.has_pending) { out.swap(s.pending); s.has_pending = false; return true; }
	out.clear();
	while (true)
	{
		if (s.pos == s.end)
		{
			if (s.eof) return !out.empty();
			int got = gzread(s.fp, s.buf.data(), (unsigned)s.buf.size());
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

int ReadSource::fill(ReadBatch& b, int max_reads, bool pair_end)
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
		b.seq.insert(b.seq.end(), seq.begin(), seq.end()); b.seq_off.push_back(b.seq.size());
		if (fastq) { qual.resize(seq.size(), ' '); b.qual.insert(b.qual.end(), qual.begin(), qual.end()); }
		b.names.insert(b.names.end(), name.begin(), name.end()); b.name_off.push_back((uint32_t)b.names.size());
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
