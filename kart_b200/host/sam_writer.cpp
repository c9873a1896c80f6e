// SAM text, byte-compatible with the reference's sprintf formats (src/Mapping.cpp:186,214,218,234,257,259,280,303 ; header :664-675).
#include "kart_host.h"
#include <string.h>
#include <algorithm>

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

static inline char* put_num(char* w, long long v)
{
	char b[24]; int n = 0; bool neg = v < 0; unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
	do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (neg) *w++ = '-';
	while (n) *w++ = b[--n];
	return w;
}
static inline char* put_str(char* w, const char* s, size_t n) { memcpy(w, s, n); return w + n; }
#define PUT_LIT(w, lit) put_str((w), (lit), sizeof(lit) - 1)

static char g_comp_tab[256]; static bool g_comp_tab_init = false;
static void comp_tab_init() { if (!g_comp_tab_init) { for (int c = 0; c < 256; c++) g_comp_tab[c] = comp_base((char)c); g_comp_tab_init = true; } }

// stored_fwd: orientation in which the read is held in the batch (mate 2 of a pair is held reverse-complemented)
void sam_read_line(HBuf<char>& o, const HostIndex& idx, const ReadBatch& b, int r, bool stored_fwd, const kb_aln_t& a, const uint32_t* cigar, bool fastq)
{
	if (a.kind == 2) return;
	if (!g_comp_tab_init) comp_tab_init();
	const char* name = b.names.data() + b.name_off[r]; size_t nlen = b.name_off[r + 1] - b.name_off[r];
	const char* seq = (const char*)b.seq.data() + b.seq_off[r]; size_t rlen = (size_t)(b.seq_off[r + 1] - b.seq_off[r]);
	const char* qual = fastq ? b.qual.data() + b.seq_off[r] : nullptr;
	const std::string* chr = a.kind == 1 ? &idx.chr_name[a.chr] : nullptr;
	size_t need = nlen + 2 * rlen + (a.kind == 1 ? chr->size() + (size_t)a.cig_len * 12 : 0) + 256;
	if (o.cap - o.n < need) o.reserve(std::max(o.n + need, o.cap * 2));
	char* w = o.p + o.n;
	w = put_str(w, name, nlen); *w++ = '\t'; w = put_num(w, a.flag);
	if (a.kind == 0)
	{
		w = PUT_LIT(w, "\t*\t0\t0\t*\t*\t0\t0\t"); w = put_str(w, seq, rlen); *w++ = '\t';
		if (fastq) { size_t q = strnlen(qual, rlen); w = put_str(w, qual, q); } else *w++ = '*';
		w = PUT_LIT(w, "\tAS:i:0\tXS:i:0\n");
		o.n = (size_t)(w - o.p);
		return;
	}
	static const char ops[] = "MIDNSHP=X";
	*w++ = '\t'; w = put_str(w, chr->data(), chr->size()); *w++ = '\t'; w = put_num(w, a.pos); *w++ = '\t'; w = put_num(w, a.mapq); *w++ = '\t';
	for (int k = 0; k < a.cig_len; k++) { uint32_t e = cigar[a.cig_off + k]; w = put_num(w, e >> 4); *w++ = ops[e & 15]; }
	if (a.mate_pos >= 0) { w = PUT_LIT(w, "\t=\t"); w = put_num(w, a.mate_pos); *w++ = '\t'; w = put_num(w, a.tlen); *w++ = '\t'; }
	else w = PUT_LIT(w, "\t*\t0\t0\t");
	bool as_is = (a.fwd != 0) == stored_fwd;
	if (as_is) w = put_str(w, seq, rlen);
	else { for (size_t i = 0; i < rlen; i++) w[i] = g_comp_tab[(unsigned char)seq[rlen - 1 - i]]; w += rlen; }
	*w++ = '\t';
	if (!fastq) *w++ = '*';
	else
	{
		size_t q = strnlen(qual, rlen);
		if (as_is) w = put_str(w, qual, q);
		else { for (size_t i = 0; i < q; i++) w[i] = qual[q - 1 - i]; w += q; }
	}
	w = PUT_LIT(w, "\tNM:i:"); w = put_num(w, (long long)rlen - a.score); w = PUT_LIT(w, "\tAS:i:"); w = put_num(w, a.score); w = PUT_LIT(w, "\tXS:i:"); w = put_num(w, a.sub_score); *w++ = '\n';
	o.n = (size_t)(w - o.p);
}
