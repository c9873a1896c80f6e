// SAM text, byte-compatible with the reference's sprintf formats (src/Mapping.cpp:186,214,218,234,257,259,280,303 ; header :664-675).
#include "kart_host.h"
#include <string.h>

static inline void put_int(std::string& o, long long v)
{
	char b[24]; int n = 0; bool neg = v < 0; unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
	do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (neg) o += '-';
	while (n) o += b[--n];
}

void sam_header(std::string& out, const HostIndex& idx)
{
	out += "@PG\tID:kart\tPN:Kart\tVN:2.5.6\n";
	for (size_t i = 0; i < idx.chr_name.size(); i++) { out += "@SQ\tSN:"; out += idx.chr_name[i]; out += "\tLN:"; put_int(out, idx.chr_len[i]); out += '\n'; }
}

static inline char comp_base(char c)
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
}

// stored_fwd: orientation in which the read is held in the batch (mate 2 of a pair is held reverse-complemented)
void sam_read_line(std::string& o, const HostIndex& idx, const ReadBatch& b, int r, bool stored_fwd, const kb_aln_t& a, const uint32_t* cigar, bool fastq)
{
	if (a.kind == 2) return;
	const char* name = b.names.data() + b.name_off[r]; size_t nlen = b.name_off[r + 1] - b.name_off[r];
	const char* seq = (const char*)b.seq.data() + b.seq_off[r]; size_t rlen = (size_t)(b.seq_off[r + 1] - b.seq_off[r]);
	const char* qual = fastq ? b.qual.data() + b.seq_off[r] : nullptr;
	o.append(name, nlen); o += '\t'; put_int(o, a.flag);
	if (a.kind == 0)
	{
		o += "\t*\t0\t0\t*\t*\t0\t0\t"; o.append(seq, rlen); o += '\t';
		if (fastq) { size_t q = strnlen(qual, rlen); o.append(qual, q); } else o += '*';
		o += "\tAS:i:0\tXS:i:0\n";
		return;
	}
	static const char ops[] = "MIDNSHP=X";
	o += '\t'; o += idx.chr_name[a.chr]; o += '\t'; put_int(o, a.pos); o += '\t'; put_int(o, a.mapq); o += '\t';
	for (int k = 0; k < a.cig_len; k++) { uint32_t e = cigar[a.cig_off + k]; put_int(o, e >> 4); o += ops[e & 15]; }
	if (a.mate_pos >= 0) { o += "\t=\t"; put_int(o, a.mate_pos); o += '\t'; put_int(o, a.tlen); o += '\t'; }
	else o += "\t*\t0\t0\t";
	bool as_is = (a.fwd != 0) == stored_fwd;
	if (as_is) o.append(seq, rlen);
	else { size_t at = o.size(); o.resize(at + rlen); for (size_t i = 0; i < rlen; i++) o[at + i] = comp_base(seq[rlen - 1 - i]); }
	o += '\t';
	if (!fastq) o += '*';
	else
	{
		size_t q = strnlen(qual, rlen);
		if (as_is) o.append(qual, q);
		else { size_t at = o.size(); o.resize(at + q); for (size_t i = 0; i < q; i++) o[at + i] = qual[q - 1 - i]; }
	}
	o += "\tNM:i:"; put_int(o, (long long)rlen - a.score); o += "\tAS:i:"; put_int(o, a.score); o += "\tXS:i:"; put_int(o, a.sub_score); o += '\n';
}
