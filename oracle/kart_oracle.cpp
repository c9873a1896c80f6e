// TEST INFRASTRUCTURE ONLY -- see kart_oracle.h. CPU restatement of Kart v2.5.6's per-read hot path,
// written from the behaviour of the reference (citations are file:line under /root/reference/src).
// Plain scalar C++; no part of this file is compiled into, or called by, the product.
#include "kart_oracle.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

// ------------------------------------------------------------------------------------------------
// Index + reference (file formats: bwt_index.cpp:16-36,103-122,38-90,230-259)
// ------------------------------------------------------------------------------------------------
struct Chrom { std::string name; i64 fwd_start, rev_start, len; };
struct Index
{
	u64 primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0;
	std::vector<u32> bwt;          // interleaved Occ/BWT blocks: 16 words per 128 symbols
	std::vector<u64> sa; int sa_intv = 32;
	i64 G = 0, G2 = 0;              // GenomeSize, TwoGenomeSize
	std::vector<Chrom> chr;
	std::vector<std::pair<i64, int> > ends;   // ChrLocMap: last coordinate of each chromosome on each strand -> idx (sorted)
	std::string text;               // RefSequence: forward + reverse complement, upper-case ACGT
	bool loaded = false;
};
struct Params { bool pacbio = false, multihit = false; int max_gaps = 5, min_seed = 13, max_insert = 1500; };

static Index g_ix;
static Params g_pm;
static std::atomic<u64> g_cnt[7];

static unsigned char nt4(unsigned char c)   // nst_nt4_table, BWT_Index/bntseq.c:40-57
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

static bool read_all(const std::string& fn, std::vector<unsigned char>& buf)
{
	FILE* fp = fopen(fn.c_str(), "rb"); if (!fp) return false;
	fseek(fp, 0, SEEK_END); long n = ftell(fp); fseek(fp, 0, SEEK_SET);
	buf.resize(n); size_t got = n ? fread(buf.data(), 1, n, fp) : 0; fclose(fp);
	return (long)got == n;
}

static int chr_lookup(i64 pos)   // index into ends of the first key >= pos, or -1 (map::lower_bound == end)
{
	const std::vector<std::pair<i64, int> >& e = g_ix.ends;
	size_t lo = 0, hi = e.size();
	while (lo < hi) { size_t mid = (lo + hi) / 2; if (e[mid].first < pos) lo = mid + 1; else hi = mid; }
	return lo == e.size() ? -1 : (int)lo;
}

static int load_index(const std::string& prefix)
{
	Index& ix = g_ix; ix = Index();
	std::vector<unsigned char> b;
	if (!read_all(prefix + ".bwt", b) || b.size() < 40) return -1;
	memcpy(&ix.primary, b.data(), 8); memcpy(&ix.L2[1], b.data() + 8, 32); ix.L2[0] = 0; ix.seq_len = ix.L2[4];
	ix.bwt.resize((b.size() - 40) / 4); memcpy(ix.bwt.data(), b.data() + 40, ix.bwt.size() * 4);
	if (!read_all(prefix + ".sa", b) || b.size() < 56) return -2;
	u64 intv; memcpy(&intv, b.data() + 40, 8); ix.sa_intv = (int)intv;
	u64 n_sa = (ix.seq_len + ix.sa_intv) / ix.sa_intv;
	ix.sa.assign(n_sa, 0); ix.sa[0] = (u64)-1;
	memcpy(ix.sa.data() + 1, b.data() + 56, std::min<size_t>((n_sa - 1) * 8, b.size() - 56));
	// .ann : "l_pac n_seqs seed" then per sequence "gi name [anno]" / "offset len n_ambs"
	FILE* fp = fopen((prefix + ".ann").c_str(), "r"); if (!fp) return -3;
	long long lpac; int nseq; unsigned seed;
	if (fscanf(fp, "%lld%d%u", &lpac, &nseq, &seed) != 3) { fclose(fp); return -3; }
	ix.G = lpac; ix.G2 = lpac * 2; ix.chr.resize(nseq);
	i64 total = 0;
	for (int i = 0; i < nseq; i++)
	{
		unsigned gi; char name[1024]; long long off; int len, namb; int c;
		if (fscanf(fp, "%u%1023s", &gi, name) != 2) { fclose(fp); return -3; }
		while ((c = fgetc(fp)) != '\n' && c != EOF) {}
		if (fscanf(fp, "%lld%d%d", &off, &len, &namb) != 3) { fclose(fp); return -3; }
		ix.chr[i].name = name; ix.chr[i].len = len; ix.chr[i].fwd_start = total; total += len; ix.chr[i].rev_start = ix.G2 - total;
		ix.ends.push_back(std::make_pair(ix.chr[i].fwd_start + len - 1, i));
		ix.ends.push_back(std::make_pair(ix.chr[i].rev_start + len - 1, i));
	}
	fclose(fp);
	std::sort(ix.ends.begin(), ix.ends.end());
	if (!read_all(prefix + ".pac", b) || (i64)b.size() < ix.G / 4) return -4;
	b.resize(ix.G / 4 + 2, 0);
	ix.text.assign(ix.G2, 'N');
	static const char fw[4] = {'A', 'C', 'G', 'T'}, rv[4] = {'T', 'G', 'C', 'A'};
	for (i64 p = 0; p < ix.G; p++) { int c = b[p >> 2] >> ((~p & 3) << 1) & 3; ix.text[p] = fw[c]; ix.text[ix.G2 - 1 - p] = rv[c]; }
	for (g_pm.min_seed = 13; g_pm.min_seed < 16; g_pm.min_seed++) if ((double)ix.G2 < pow(4, g_pm.min_seed)) break;   // Mapping.cpp:645
	ix.loaded = true;
	return 0;
}

// ------------------------------------------------------------------------------------------------
// FM-index primitives (bwt_search.cpp:44-138). Counting is done symbol by symbol, which is the
// definition the reference's LUT / SWAR popcount shortcuts implement.
// ------------------------------------------------------------------------------------------------
static inline int bwt_sym(u64 r) { const u32* blk = &g_ix.bwt[(r >> 7) << 4]; return blk[8 + ((r & 127) >> 4)] >> ((~r & 15) << 1) & 3; }

static void occ4(u64 k, u64 cnt[4])   // number of each base in BWT rows [0..k] ($ skipped); k == -1 -> zeros
{
	if (k == (u64)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= g_ix.primary);
	const u32* blk = &g_ix.bwt[(k >> 7) << 4];
	memcpy(cnt, blk, 32);
	u64 base = k & ~(u64)127;
	for (u64 r = base; r <= k; r++) cnt[bwt_sym(r)]++;
}

static u64 occ1(u64 k, int c)   // bwt_occ (:44)
{
	if (k == g_ix.seq_len) return g_ix.L2[c + 1] - g_ix.L2[c];
	if (k == (u64)-1) return 0;
	u64 cnt[4]; occ4(k, cnt); return cnt[c];
}

static u64 inv_psi(u64 k)   // bwt_invPsi (:120): note '>' here vs '>=' inside occ
{
	u64 x = k - (k > g_ix.primary);
	int c = bwt_sym(x);
	u64 r = g_ix.L2[c] + occ1(k, c);
	return k == g_ix.primary ? 0 : r;
}

static u64 sa_locate(u64 k)   // bwt_sa (:128)
{
	u64 steps = 0, mask = g_ix.sa_intv - 1;
	while (k & mask) { steps++; k = inv_psi(k); }
	g_cnt[3]++; g_cnt[4] += steps;
	return steps + g_ix.sa[k / g_ix.sa_intv];
}

struct SearchHit { int len, freq; u64 x0, x2; std::vector<u64> loc; };

static SearchHit fm_search(const unsigned char* q, int start, int stop)   // BWT_Search (:140)
{
	const Index& ix = g_ix; SearchHit h;
	int p = q[start], pos;
	u64 x0 = ix.L2[p] + 1, x1 = ix.L2[3 - p] + 1, x2 = ix.L2[p + 1] - ix.L2[p];
	g_cnt[0]++;
	for (pos = start + 1; pos < stop; pos++)
	{
		if (q[pos] > 3) break;
		u64 tk[4], tl[4], k = x1 - 1, l = x1 - 1 + x2;
		occ4(k, tk); occ4(l, tl);
		{   // work accounting: distinct 64-byte blocks touched (bwt_search.cpp:92)
			u64 kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
			g_cnt[1]++; g_cnt[2] += (k == (u64)-1 || l == (u64)-1 || (kk >> 7) != (ll >> 7)) ? ((k == (u64)-1 || l == (u64)-1) ? 1 : 2) : 1;
		}
		u64 nx1[4], nx2[4], nx0[4];
		for (int c = 0; c < 4; c++) { nx1[c] = ix.L2[c] + 1 + tk[c]; nx2[c] = tl[c] - tk[c]; }
		nx0[3] = x0 + ((x1 <= ix.primary && x1 + x2 - 1 >= ix.primary) ? 1 : 0);
		nx0[2] = nx0[3] + nx2[3]; nx0[1] = nx0[2] + nx2[2]; nx0[0] = nx0[1] + nx2[1];
		int c = 3 - q[pos];
		if (nx2[c] == 0) break;
		x0 = nx0[c]; x1 = nx1[c]; x2 = nx2[c];
	}
	h.len = pos - start; h.freq = 0; h.x0 = x0; h.x2 = x2;
	if (h.len >= g_pm.min_seed && (int)x2 <= 50)
	{
		h.freq = (int)x2;
		for (int i = 0; i < h.freq; i++) h.loc.push_back(sa_locate(x0 + i));
	}
	return h;
}

// ------------------------------------------------------------------------------------------------
// Seeds, candidates (AlignmentCandidates.cpp:49-224)
// ------------------------------------------------------------------------------------------------
struct Seg { bool simple; int rPos, rLen, gLen; i64 gPos, diff; };
struct Cand { int score; i64 diff; int mate; std::vector<Seg> segs; };

static bool by_diff(const Seg& a, const Seg& b) { return a.diff == b.diff ? a.rPos < b.rPos : a.diff < b.diff; }
static bool by_gpos(const Seg& a, const Seg& b) { return a.gPos == b.gPos ? a.rPos < b.rPos : a.gPos < b.gPos; }

static void encode(const char* s, int n, std::vector<unsigned char>& q) { q.resize(n + 1); for (int i = 0; i < n; i++) q[i] = nt4(s[i]); q[n] = 4; }

static std::vector<Seg> seeds_fast(int rlen, const unsigned char* q)   // :49
{
	std::vector<Seg> v; int pos = 0, end = rlen - g_pm.min_seed;
	while (pos < end)
	{
		if (q[pos] > 3) { pos++; continue; }
		SearchHit h = fm_search(q, pos, rlen);
		for (int i = 0; i < h.freq; i++) { Seg s; s.simple = true; s.rPos = pos; s.rLen = s.gLen = h.len; s.gPos = (i64)h.loc[i]; s.diff = s.gPos - pos; v.push_back(s); }
		pos += h.len + 1;
	}
	std::sort(v.begin(), v.end(), by_diff);
	return v;
}

static std::vector<Seg> seeds_sensitive(int rlen, const unsigned char* q)   // :132
{
	std::vector<Seg> v; int pos = 0, stop = 30, end = rlen - g_pm.min_seed;
	while (pos < end)
	{
		if (q[pos] > 3) { pos++; stop++; continue; }
		SearchHit h = fm_search(q, pos, stop);
		if (h.freq > 0)
		{
			for (int i = 0; i < h.freq; i++) { Seg s; s.simple = true; s.rPos = pos; s.rLen = s.gLen = h.len; s.gPos = (i64)h.loc[i]; s.diff = s.gPos - pos; v.push_back(s); }
			pos += h.len; stop += h.len;
		}
		else { pos += g_pm.min_seed; stop += g_pm.min_seed; }
		if (stop > rlen) stop = rlen;
	}
	std::sort(v.begin(), v.end(), by_gpos);
	return v;
}

static std::vector<Cand> cands_illumina(int rlen, const std::vector<Seg>& sv)   // :82
{
	std::vector<Cand> out; int thr = (int)(rlen * 0.2); if (thr > 50) thr = 50;
	int n = (int)sv.size(), i = 0;
	while (i < n && sv[i].diff < 0) i++;
	while (i < n)
	{
		int score = sv[i].rLen, j = i, k; int e = chr_lookup(sv[i].gPos); i64 bound = g_ix.ends[e].first;
		for (k = i + 1; k < n; k++)
		{
			if (sv[k].gPos > bound || sv[k].diff - sv[j].diff > g_pm.max_gaps) break;
			score += sv[k].rLen; j = k;
		}
		if (score > thr)
		{
			Cand c; c.score = score; c.mate = -1; c.segs.assign(sv.begin() + i, sv.begin() + k);
			if (score - 50 > thr) thr = score - 50;
			c.diff = c.segs[0].diff < 0 ? 0 : c.segs[0].diff;
			std::sort(c.segs.begin(), c.segs.end(), by_gpos);
			out.push_back(c);
		}
		i = k;
	}
	return out;
}

static std::vector<Cand> cands_pacbio(int rlen, const std::vector<Seg>& sv)   // :171
{
	(void)rlen; std::vector<Cand> out; int n = (int)sv.size(); if (n == 0) return out;
	std::vector<char> taken(n, 0); int thr = 0, i = 0;
	while (i < n && sv[i].diff < 0) i++;
	for (; i < n; i++)
	{
		if (taken[i]) continue;
		Cand c; c.score = sv[i].rLen; c.mate = -1; taken[i] = 1; c.segs.push_back(sv[i]);
		int j = i;
		for (int k = i + 1; k < n; k++)
		{
			if (taken[k]) continue;
			i64 d = sv[k].diff - sv[j].diff; if (d < 0) d = -d;
			if (d < 300)
			{
				if (sv[k].rPos > sv[j].rPos) { c.score += sv[k].rLen; c.segs.push_back(sv[k]); taken[k] = 1; j = k; }
			}
			else if (sv[k].gPos - sv[j].gPos > 1000) break;
		}
		if (c.score >= thr) { thr = c.score; c.diff = sv[i].diff < 0 ? 0 : sv[i].diff; out.push_back(c); }
	}
	return out;
}

// ------------------------------------------------------------------------------------------------
// Normal-pair identification (AlignmentCandidates.cpp:226-490)
// ------------------------------------------------------------------------------------------------
static void drop_empty(std::vector<Seg>& v) { size_t w = 0; for (size_t i = 0; i < v.size(); i++) if (v[i].rLen != 0) v[w++] = v[i]; v.resize(w); }

static void drop_shared_rpos(std::vector<Seg>& v)   // RemoveTandemRepeatSeeds :235
{
	int n = (int)v.size(); if (n < 2) return;
	std::vector<std::pair<int, int> > o(n);
	for (int i = 0; i < n; i++) o[i] = std::make_pair(v[i].rPos, i);
	std::sort(o.begin(), o.end());
	bool any = false;
	for (int i = 0; i < n;)
	{
		int j = i + 1; while (j < n && o[j].first == o[i].first) j++;
		if (j - i > 1) { any = true; for (int k = i; k < j; k++) v[o[k].second].rLen = v[o[k].second].gLen = 0; }
		i = j;
	}
	if (any) drop_empty(v);
}

static void drop_translocated(std::vector<Seg>& v)   // RemoveTranslocatedSeeds :273 (+ IdentifyTranslocationRange :262)
{
	int n = (int)v.size(); if (n < 2) return;
	std::vector<std::pair<int, int> > o(n);
	for (int i = 0; i < n; i++) o[i] = std::make_pair(v[i].rPos, i);
	std::sort(o.begin(), o.end());   // rPos are distinct here, so the order is unique
	bool any = false;
	for (int i = 0; i < n; i++)
	{
		if (o[i].first == v[i].rPos) continue;
		any = true;
		int hi = o[i].second;
		for (int j = i + 1; j <= hi; j++) if (o[j].second > hi) hi = o[j].second;
		int s1 = 0, s2 = 0;
		for (int k = i; k <= hi; k++) { if (k < o[k].second) s1 += v[o[k].second].rLen; else s2 += v[o[k].second].rLen; }
		for (int k = i; k <= hi; k++)
		{
			bool kill = (s1 > s2) ? (k > o[k].second) : (k < o[k].second);
			if (kill) v[o[k].second].rLen = v[o[k].second].gLen = 0;
		}
		i = hi;
	}
	if (any) drop_empty(v);
}

static bool resolve_overlap(Seg& a, Seg& b)   // CheckSeedOverlapping :323 ; false = a yielded
{
	bool master = true; int ov;
	if ((ov = a.rPos + a.rLen - b.rPos) > 0)
	{
		if (a.rLen < b.rLen) { master = false; if (a.rLen > ov) a.gLen = (a.rLen -= ov); else a.rLen = a.gLen = 0; }
		else if (b.rLen > ov) { b.rPos += ov; b.gPos += ov; b.gLen = (b.rLen -= ov); }
		else b.rLen = b.gLen = 0;
	}
	if (a.rLen > 0 && b.rLen > 0 && (ov = (int)(a.gPos + a.gLen - b.gPos)) > 0)
	{
		if (a.gLen < b.gLen) { master = false; if (a.rLen > ov) a.gLen = (a.rLen -= ov); else a.rLen = a.gLen = 0; }
		else if (b.rLen > ov) { b.rPos += ov; b.gPos += ov; b.gLen = (b.rLen -= ov); }
		else b.rLen = b.gLen = 0;
	}
	return master;
}

static void trim_overlaps(std::vector<Seg>& v)   // CheckOverlappingSeeds :382
{
	int n = (int)v.size(); if (n < 2) return;
	bool any = false;
	for (int i = 0; i < n;)
	{
		if (v[i].rLen > 0)
		{
			int rEnd = v[i].rPos + v[i].rLen - 1; i64 gEnd = v[i].gPos + v[i].gLen - 1;   // deliberately not refreshed inside the loop
			for (int j = i + 1; j < n; j++)
			{
				if (v[j].rLen == 0) continue;
				if (rEnd < v[j].rPos && gEnd < v[j].gPos) break;
				if (!resolve_overlap(v[i], v[j])) break;
			}
			if (v[i].rLen == 0)
			{
				any = true;
				int p = i - 1; while (p > 0 && v[p].rLen == 0) p--;   // LocateThePreviousSeedIdx :375
				i = p < 0 ? 0 : p;
			}
			else i++;
		}
		else { any = true; i++; }
	}
	if (any) drop_empty(v);
}

static void fill_pairs(int rlen, int glen, std::vector<Seg>& v)   // IdentifyNormalPairs :420
{
	Seg g; g.simple = false; g.diff = 0;
	if (v.size() > 1)
	{
		drop_shared_rpos(v); drop_translocated(v); trim_overlaps(v);
		int n = (int)v.size();
		for (int i = 0, j = 1; j < n; i++, j++)
		{
			int rg = v[j].rPos - (v[i].rPos + v[i].rLen); if (rg < 0) rg = 0;
			int gg = (int)(v[j].gPos - (v[i].gPos + v[i].gLen)); if (gg < 0) gg = 0;
			if (rg > 0 || gg > 0)
			{
				g.rPos = v[i].rPos + v[i].rLen; g.gPos = v[i].gPos + v[i].gLen; g.diff = g.gPos - g.rPos; g.rLen = rg; g.gLen = gg;
				v.push_back(g);
			}
		}
		if ((int)v.size() > n) std::inplace_merge(v.begin(), v.begin() + n, v.end(), by_gpos);
	}
	if (v.size() > 0)
	{
		int rg = v[0].rPos > 0 ? v[0].rPos : 0;
		int gg = glen > 0 ? (int)v[0].gPos : rg;
		if (rg > 0 || gg > 0)
		{
			g.rPos = 0; g.gPos = v[0].gPos - gg; if (g.gPos < 0) g.gPos = 0;   // (:464 adds zero to gGaps)
			g.diff = g.gPos; g.simple = false; g.rLen = rg; g.gLen = gg;
			v.insert(v.begin(), g);
		}
		const Seg& t = v.back();
		rg = rlen - (t.rPos + t.rLen);
		gg = glen > 0 ? glen - (int)(t.gPos + t.gLen) : rg;
		if (rg > 0 || gg > 0)
		{
			g.simple = false; g.rPos = t.rPos + t.rLen; g.gPos = t.gPos + t.gLen; g.rLen = rg; g.gLen = gg;   // diff left stale (:479-484)
			v.push_back(g);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// 8-mer analysis (KmerAnalysis.cpp)
// ------------------------------------------------------------------------------------------------
struct Kmer { u32 wid, pos; };
struct KPair { int diff; u32 rPos, gPos; };

static u32 fresh_kmer(const char* s, u32 pos) { u32 id = 0; for (u32 i = pos; i < pos + 8; i++) id = (id << 2) + nt4(s[i]); return id; }   // :25

static std::vector<Kmer> kmers_of(int len, const char* s)   // CreateKmerVecFromReadSeq :56
{
	std::vector<Kmer> v; u32 count = 0, head, tail = 0, n = (u32)(len < 0 ? 0 : len);
	while (count < 8 && tail < n) { if (s[tail++] != 'N') count++; else count = 0; }
	if (count != 8) return v;
	Kmer w; w.pos = head = tail - 8; w.wid = fresh_kmer(s, head); v.push_back(w);
	for (head += 1; tail < n; head++, tail++)
	{
		if (s[tail] != 'N') { w.pos = head; w.wid = ((w.wid & 0x3FFF) << 2) + nt4(s[tail]); v.push_back(w); }
		else
		{
			count = 0; tail++;
			while (count < 8 && tail < n) { if (s[tail++] != 'N') count++; else count = 0; }
			if (count != 8) break;
			w.pos = head = tail - 8; w.wid = fresh_kmer(s, head); v.push_back(w);
		}
	}
	std::stable_sort(v.begin(), v.end(), [](const Kmer& a, const Kmer& b) { return a.wid < b.wid; });
	return v;
}

static std::vector<KPair> common_kmers(int max_shift, const std::vector<Kmer>& a, const std::vector<Kmer>& b)   // :104
{
	std::vector<KPair> out;
	for (size_t i = 0; i < a.size(); i++)
	{
		size_t lo = std::lower_bound(b.begin(), b.end(), a[i], [](const Kmer& x, const Kmer& y) { return x.wid < y.wid; }) - b.begin();
		for (; lo < b.size() && b[lo].wid == a[i].wid; lo++)
		{
			u32 d = b[lo].pos >= a[i].pos ? b[lo].pos - a[i].pos : a[i].pos - b[lo].pos;
			if (d < (u32)max_shift) { KPair p; p.rPos = a[i].pos; p.gPos = b[lo].pos; p.diff = (int)(p.gPos - p.rPos); out.push_back(p); }
		}
	}
	std::sort(out.begin(), out.end(), [](const KPair& x, const KPair& y) { return x.diff == y.diff ? x.rPos < y.rPos : x.diff < y.diff; });
	return out;
}

static std::vector<Seg> runs_to_pairs(int min_len, const std::vector<KPair>& kp)   // GenerateSimplePairsFromCommonKmers :132
{
	std::vector<Seg> out; int n = (int)kp.size();
	for (int i = 0; i < n;)
	{
		int j = i + 1; u32 next = kp[i].rPos + 1;
		for (; j < n; j++) { if (kp[j].rPos != next || kp[j].diff != kp[i].diff) break; next++; }
		int l = 8 + (j - 1 - i);
		if (l >= min_len) { Seg s; s.simple = true; s.rPos = (int)kp[i].rPos; s.gPos = kp[i].gPos; s.diff = kp[i].diff; s.rLen = s.gLen = l; out.push_back(s); }
		i = j;
	}
	return out;
}

static std::vector<Seg> fragment_pairs(int max_dist, int l1, const char* f1, int l2, const char* f2)   // :164
{
	std::vector<Kmer> a = kmers_of(l1, f1), b = kmers_of(l2, f2);
	std::vector<Seg> v = runs_to_pairs(8, common_kmers(max_dist, a, b));
	std::sort(v.begin(), v.end(), by_gpos);
	return v;
}

// ------------------------------------------------------------------------------------------------
// Needleman-Wunsch (nw_alignment.cpp:18-80) as the exact integer recurrence: every reference float
// is a multiple of 0.5, so all values here are 2x the reference's.
// ------------------------------------------------------------------------------------------------
static void nw_align(std::string& s1, std::string& s2)
{
	int m = (int)s1.size(), n = (int)s2.size(), W = n + 1;
	g_cnt[5]++; g_cnt[6] += (u64)m * n;
	std::vector<int> R((size_t)(m + 1) * W), T((size_t)(m + 1) * W), S((size_t)(m + 1) * W);
	const int NEG = -131072;
	R[0] = T[0] = S[0] = 0;
	for (int i = 1; i <= m; i++) { R[(size_t)i * W] = NEG; S[(size_t)i * W] = T[(size_t)i * W] = -2 - i; }
	for (int j = 1; j <= n; j++) { T[j] = NEG; S[j] = R[j] = -2 - j; }
	for (int i = 1; i <= m; i++)
	{
		unsigned char a = nt4(s1[i - 1]);
		for (int j = 1; j <= n; j++)
		{
			size_t c = (size_t)i * W + j;
			int r = std::max(R[c - 1] - 1, S[c - 1] - 3), t = std::max(T[c - W] - 1, S[c - W] - 3);
			int d = S[c - W - 1] + (a == nt4(s2[j - 1]) ? 3 : -3);
			R[c] = r; T[c] = t; S[c] = std::max(d, std::max(r, t));
		}
	}
	int i = m, j = n;
	while (i > 0 || j > 0)
	{
		size_t c = (size_t)i * W + j;
		if (S[c] == R[c]) { s1.insert(i, 1, '-'); j--; }
		else if (S[c] == T[c]) { s2.insert(j, 1, '-'); i--; }
		else { i--; j--; }
	}
}

// ------------------------------------------------------------------------------------------------
// Fragment-pair processing (tools.cpp:40-104,142-397)
// ------------------------------------------------------------------------------------------------
typedef std::vector<std::pair<int, char> > Cigar;

static int mismatches(const std::string& a, const std::string& b) { int c = 0; for (size_t i = 0; i < a.size(); i++) if (a[i] != b[i]) c++; return c; }

static int push_alignment(const std::string& a, const std::string& b, Cigar& cg)   // AddNewCigarElements :49
{
	char st = '*'; int c = 0, score = 0;
	for (size_t i = 0; i < a.size(); i++)
	{
		char now = a[i] == '-' ? 'D' : (b[i] == '-' ? 'I' : 'M');
		if (now == 'M' && a[i] == b[i]) score++;
		if (now == st) c++;
		else { if (c > 0) cg.push_back(std::make_pair(c, st)); c = 1; st = now; }
	}
	if (c > 0) cg.push_back(std::make_pair(c, st));
	return score;
}

static void align_fragments(std::string& f1, std::string& f2)   // GenerateNormalPairAlignment :142
{
	int rl = (int)f1.size(), gl = (int)f2.size();
	if (rl > 30 && gl > 30)
	{
		int shift;
		if (g_pm.pacbio) { shift = rl > gl ? (int)(rl * 0.2) : (int)(gl * 0.2); if (shift > 50) shift = 50; }
		else shift = g_pm.max_gaps;
		std::vector<Seg> part = fragment_pairs(shift, rl, f1.c_str(), gl, f2.c_str());
		if (part.size() > 0) fill_pairs(rl, gl, part);
		if (part.size() > 0)
		{
			std::string a1, a2;
			for (size_t i = 0; i < part.size(); i++)
			{
				const Seg& p = part[i];
				if (p.rLen <= 0 && p.gLen <= 0) continue;
				if (p.gLen == 0) { a1 += f1.substr(p.rPos, p.rLen); a2 += std::string(p.rLen, '-'); }
				else if (p.rLen == 0) { a1 += std::string(p.gLen, '-'); a2 += f2.substr(p.gPos, p.gLen); }
				else
				{
					std::string x = f1.substr(p.rPos, p.rLen), y = f2.substr(p.gPos, p.gLen);
					if (!(p.rLen == 1 && p.gLen == 1) && !p.simple)
					{
						if (g_pm.pacbio && (p.rLen > 300 || p.gLen > 300)) align_fragments(x, y);
						else nw_align(x, y);
					}
					a1 += x; a2 += y;
				}
			}
			f1 = a1; f2 = a2;
			return;
		}
	}
	nw_align(f1, f2);
}

static bool quick_match(const std::string& a, const std::string& b, int& n)   // tools.cpp:240,301,352
{
	if (a.size() != b.size()) return false;
	n = mismatches(a, b);
	return n <= 2 && n <= (int)(a.size() * 0.2);
}

static bool local_quality_ok(const std::string& a, const std::string& b)   // CheckLocalAlignmentQuality :255
{
	int type = -1, runs = 0, n = 0, mis = 0;
	for (size_t i = 0; i < a.size(); i++)
	{
		int t = a[i] == '-' ? 0 : (b[i] == '-' ? 1 : 2);
		if (t == 2) { n++; if (a[i] != b[i]) mis++; }
		if (t != type) { type = t; runs++; }
	}
	return !(runs >= 4 || (mis >= 3 && mis >= (int)(n * 0.3)));
}

static std::string ref_slice(i64 pos, int len)   // RefSequence + pos ; out-of-range bytes (UB in the reference) read as 'N'
{
	std::string s(len > 0 ? len : 0, 'N');
	for (int i = 0; i < len; i++) { i64 p = pos + i; if (p >= 0 && p < g_ix.G2) s[i] = g_ix.text[p]; }
	return s;
}

static int do_middle(const char* seq, Seg& sp, Cigar& cg)   // ProcessNormalSequencePair :225
{
	if (sp.rLen == 0 || sp.gLen == 0)
	{
		if (sp.rLen > 0) cg.push_back(std::make_pair(sp.rLen, 'I')); else if (sp.gLen > 0) cg.push_back(std::make_pair(sp.gLen, 'D'));
		return 0;
	}
	std::string f1(seq + sp.rPos, sp.rLen), f2 = ref_slice(sp.gPos, sp.gLen); int n;
	if (quick_match(f1, f2, n)) { cg.push_back(std::make_pair(sp.rLen, 'M')); return sp.rLen - n; }
	align_fragments(f1, f2);
	return push_alignment(f1, f2, cg);
}

static int do_head(const char* seq, Seg& sp, Cigar& cg)   // ProcessHeadSequencePair :292
{
	std::string f1(seq + sp.rPos, sp.rLen), f2 = ref_slice(sp.gPos, sp.gLen); int n;
	if (!g_pm.pacbio && quick_match(f1, f2, n)) { cg.push_back(std::make_pair(sp.rLen, 'M')); return sp.rLen - n; }
	if (!g_pm.pacbio && sp.rLen > 50) { cg.push_back(std::make_pair(sp.rLen, 'S')); return 0; }
	align_fragments(f1, f2);
	if (!local_quality_ok(f1, f2)) { cg.push_back(std::make_pair(sp.rLen, 'S')); return 0; }
	size_t p = 0; while (p < f1.size() && f1[p] == '-') p++;
	if (p > 0) { f1.erase(0, p); f2.erase(0, p); sp.gPos += p; sp.gLen -= (int)p; }
	p = 0; while (p < f2.size() && f2[p] == '-') p++;
	if (p > 0) { f1.erase(0, p); f2.erase(0, p); sp.rPos += (int)p; sp.rLen -= (int)p; cg.push_back(std::make_pair((int)p, 'S')); }
	return push_alignment(f1, f2, cg);
}

static int do_tail(const char* seq, Seg& sp, Cigar& cg)   // ProcessTailSequencePair :344
{
	std::string f1(seq + sp.rPos, sp.rLen), f2 = ref_slice(sp.gPos, sp.gLen); int n;
	if (!g_pm.pacbio && quick_match(f1, f2, n)) { cg.push_back(std::make_pair(sp.rLen, 'M')); return sp.rLen - n; }
	if (!g_pm.pacbio && sp.rLen > 100) { cg.push_back(std::make_pair(sp.rLen, 'S')); return 0; }
	align_fragments(f1, f2);
	if (!local_quality_ok(f1, f2)) { cg.push_back(std::make_pair(sp.rLen, 'S')); return 0; }
	int c = 0; for (int p = (int)f1.size() - 1; p >= 0 && f1[p] == '-'; p--) c++;
	if (c > 0) { f1.resize(f1.size() - c); f2.resize(f2.size() - c); sp.gLen -= c; }
	c = 0; for (int p = (int)f2.size() - 1; p >= 0 && f2[p] == '-'; p--) c++;
	if (c > 0) { f1.resize(f1.size() - c); f2.resize(f2.size() - c); sp.rLen -= c; }
	int score = push_alignment(f1, f2, cg);
	if (c > 0) cg.push_back(std::make_pair(c, 'S'));
	return score;
}

// ------------------------------------------------------------------------------------------------
// Reports, coordinates (AlignmentCandidates.cpp:492-745)
// ------------------------------------------------------------------------------------------------
struct Report { int aln = 0, flag = 0, mate = -1; bool fwd = true; int chr = 0; i64 pos = 0; std::string cigar; };
struct ReadRes { int rlen = 0, mapq = 0, score = 0, sub = 0, ncan = 0, best = 0; std::vector<Report> rep; };

static std::string cigar_text(const Cigar& cg)   // GenerateCIGAR :492
{
	std::string s; char st = 0; int c = 0; char buf[24];
	for (size_t i = 0; i < cg.size(); i++)
	{
		if (cg[i].second != st) { if (c > 0) { snprintf(buf, sizeof(buf), "%d%c", c, st); s += buf; } c = cg[i].first; st = cg[i].second; }
		else c += cg[i].first;
	}
	if (c > 0) { snprintf(buf, sizeof(buf), "%d%c", c, st); s += buf; }
	return s;
}

static void locate(bool first, i64 gpos, i64 gend, Cigar& cg, Report& r)   // GenCoordinateInfo :515
{
	const Index& ix = g_ix;
	if (gpos < ix.G)
	{
		r.fwd = first;
		if (ix.chr.size() == 1) { r.chr = 0; r.pos = gpos + 1; }
		else { int e = chr_lookup(gpos); r.chr = ix.ends[e].second; r.pos = gpos + 1 - ix.chr[r.chr].fwd_start; }
	}
	else
	{
		r.fwd = !first;
		std::reverse(cg.begin(), cg.end());
		if (ix.chr.size() == 1) { r.chr = 0; r.pos = ix.G2 - gend; }
		else { int e = chr_lookup(gpos); r.pos = ix.ends[e].first - gend + 1; r.chr = ix.ends[e].second; }
	}
	r.cigar = cigar_text(cg);
}

static bool same_chromosome(const std::vector<Seg>& v)   // CheckCoordinateValidity :582
{
	i64 a = 0, b = g_ix.G2;
	for (size_t i = 0; i < v.size(); i++) if (v[i].gLen > 0) { a = v[i].gPos; break; }
	for (size_t i = v.size(); i-- > 0;) if (v[i].gLen > 0) { b = v[i].gPos + v[i].gLen - 1; break; }
	if ((a < g_ix.G) != (b < g_ix.G)) return false;
	int ea = chr_lookup(a), eb = chr_lookup(b);
	return ea >= 0 && eb >= 0 && g_ix.ends[ea].second == g_ix.ends[eb].second;
}

static void make_reports(bool first, const char* seq, ReadRes& rd, std::vector<Cand>& cv)   // GenMappingReport :624
{
	rd.score = rd.sub = rd.best = 0;
	if ((rd.ncan = (int)cv.size()) == 0) { rd.ncan = 1; rd.rep.assign(1, Report()); return; }
	rd.rep.assign(rd.ncan, Report());
	for (int i = 0; i < rd.ncan; i++)
	{
		Report& rp = rd.rep[i]; rp.aln = 0; rp.mate = cv[i].mate;
		if (cv[i].score == 0) continue;
		if (g_pm.pacbio && rd.score > 0) { rd.sub = rd.score; continue; }
		std::vector<Seg>& sv = cv[i].segs;
		fill_pairs(rd.rlen, -1, sv);
		if (!same_chromosome(sv)) continue;
		Cigar cg; int n = (int)sv.size();
		for (int j = 0; j < n; j++)
		{
			if (sv[j].rLen == 0 && sv[j].gLen == 0) continue;
			if (sv[j].simple) { cg.push_back(std::make_pair(sv[j].rLen, 'M')); rp.aln += sv[j].rLen; continue; }
			if (j == 0)
			{
				int s = 0;
				if (sv[0].rLen > 3000) cg.push_back(std::make_pair(sv[0].rLen, 'S'));
				else { s = do_head(seq, sv[0], cg); rp.aln += s; }
				if (s == 0) { sv[0].gPos = sv[1].gPos; sv[0].gLen = 0; }
			}
			else if (j == n - 1)
			{
				int s = 0;
				if (sv[j].rLen > 3000) cg.push_back(std::make_pair(sv[j].rLen, 'S'));
				else { s = do_tail(seq, sv[j], cg); rp.aln += s; }
				if (s == 0) { sv[j].gPos = sv[j - 1].gPos + sv[j - 1].gLen; sv[j].gLen = 0; }
			}
			else rp.aln += do_middle(seq, sv[j], cg);
		}
		if (!g_pm.pacbio && cg.size() > 1)
		{
			int gp = 0; for (size_t k = 0; k < cg.size(); k++) if (cg[k].second == 'I' || cg[k].second == 'D') gp += cg[k].first;   // GapPenalty :612
			rp.aln -= gp;
			if (rp.aln <= 0) { rp.aln = 0; continue; }
		}
		if (cg.size() == 0) rp.aln = 0;
		else { locate(first, sv[0].gPos, sv[n - 1].gPos + sv[n - 1].gLen - 1, cg, rp); if (rp.pos <= 0) rp.aln = 0; }
		if (rp.aln > rd.score) { rd.best = i; rd.sub = rd.score; rd.score = rp.aln; }
		else if (rp.aln == rd.score)
		{
			rd.sub = rd.score;
			if (!g_pm.multihit && g_ix.chr[rp.chr].len > g_ix.chr[rd.rep[rd.best].chr].len) rd.best = i;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Pairing, rescue, flags, MAPQ (Mapping.cpp:49-175,317-480 ; AlignmentRescue.cpp)
// ------------------------------------------------------------------------------------------------
static void prune_candidates(std::vector<Cand>& v)   // RemoveRedundantCandidates :317
{
	if (v.size() <= 1) return;
	int s1 = 0, s2 = 0;
	for (size_t i = 0; i < v.size(); i++) if (v[i].score > s2) { if (v[i].score >= s1) { s2 = s1; s1 = v[i].score; } else s2 = v[i].score; }
	int thr = (g_pm.pacbio || s1 == s2 || s1 - s2 > 20) ? s1 : s2;
	for (size_t i = 0; i < v.size(); i++) if (v[i].score < thr) v[i].score = 0;
}

static bool pair_candidates(i64 est, std::vector<Cand>& a, std::vector<Cand>& b)   // CheckPairedAlignmentCandidates :348
{
	int n1 = (int)a.size(), n2 = (int)b.size(); bool any = false;
	if (n1 * n2 > 1000) { prune_candidates(a); prune_candidates(b); }
	for (int i = 0; i < n1; i++)
	{
		if (a[i].score == 0) continue;
		int best = -1, s = 0;
		for (int j = 0; j < n2; j++)
		{
			if (b[j].score == 0 || b[j].diff < a[i].diff) continue;
			if (b[j].diff - a[i].diff < est) { if (b[j].score > s) { best = j; s = b[j].score; } else if (b[j].score == s) best = -1; }
		}
		if (s > 0 && best != -1)
		{
			int j = best;
			if (b[j].mate == -1) { any = true; a[i].mate = j; b[j].mate = i; }
			else if (a[i].score > a[b[j].mate].score) { a[b[j].mate].mate = -1; a[i].mate = j; b[j].mate = i; }
		}
	}
	return any;
}

static void keep_mated(std::vector<Cand>& a, std::vector<Cand>& b)   // RemoveUnMatedAlignmentCandidates :402
{
	for (size_t i = 0; i < a.size(); i++)
	{
		if (a[i].mate == -1) a[i].score = 0;
		else { int j = a[i].mate; a[i].score = b[j].score = a[i].score + b[j].score; }
	}
	for (size_t j = 0; j < b.size(); j++) if (b[j].mate == -1) b[j].score = 0;
}

static int top_score(const std::vector<Cand>& v) { int s = 0; for (size_t i = 0; i < v.size(); i++) if (v[i].score > s) s = v[i].score; return s; }

static Cand rescue_in_window(i64 left, std::vector<Seg>& v)   // IdnetifyRescueCandidate, AlignmentRescue.cpp:26
{
	Cand c; c.score = 0; c.mate = -1; c.diff = 0; int n = (int)v.size();
	for (int i = 0; i < n;)
	{
		int s = v[i].rLen, j; v[i].gPos += left;
		std::vector<Seg> grp(1, v[i]);
		for (j = i + 1; j < n; j++)
		{
			if (v[j].diff - v[i].diff < g_pm.max_gaps) { v[j].gPos += left; s += v[j].rLen; grp.push_back(v[j]); }
			else break;
		}
		if (s > c.score) { c.score = s; c.diff = grp[0].diff + left; c.segs = grp; }
		i = j;
	}
	std::sort(c.segs.begin(), c.segs.end(), by_gpos);
	for (size_t i = 0; i < c.segs.size(); i++) c.segs[i].diff += left;
	return c;
}

static bool rescue_pair(int est, int l1, const char* s1, int l2, const char* s2, std::vector<Cand>& a, std::vector<Cand>& b)   // RescueUnpairedAlignment :73
{
	const Index& ix = g_ix;
	int sc1 = top_score(a), sc2 = top_score(b), strategy; bool mated = false;
	if (sc1 == 0 && sc2 == 0) return false;
	if (sc1 < (int)(l1 * 0.1) && sc2 < (int)(l2 * 0.1)) strategy = 4;
	else if (sc1 > sc2 && sc1 - sc2 > 50) strategy = 1;
	else if (sc2 > sc1 && sc2 - sc1 > 50) strategy = 2;
	else strategy = 3;
	if (est > g_pm.max_insert) est = g_pm.max_insert;
	int n1 = (int)a.size(), n2 = (int)b.size();
	if (strategy == 1 || strategy == 3)
	{
		int thr = std::max(sc1 - 30, 50);   // DetermineAnchorThreshold :14
		std::vector<Kmer> k1 = kmers_of(l2, s2);
		for (int j = n2, i = 0; i < n1; i++)
		{
			if (a[i].score < thr) continue;
			i64 left = a[i].diff, right = a[i].diff + est + l2;
			int e = chr_lookup(left); if (e < 0) continue;   // (reference dereferences end(): UB)
			int cid = ix.ends[e].second;
			if (right < ix.G && right > ix.chr[cid].fwd_start) right = ix.chr[cid].fwd_start - 1;
			else if (right >= ix.G && right > ix.chr[cid].rev_start) right = ix.chr[cid].rev_start - 1;
			int slen = (int)(right - left); if (slen < l2) continue;
			std::string win = ref_slice(left, slen);
			std::vector<Kmer> k2 = kmers_of(slen, win.c_str());
			std::vector<Seg> sp = runs_to_pairs(10, common_kmers(slen, k1, k2));
			Cand c = rescue_in_window(left, sp);
			if (c.score > sc2) { mated = true; c.mate = i; a[i].mate = j++; b.push_back(c); }
		}
	}
	if (strategy == 2 || strategy == 3)
	{
		int thr = std::max(sc2 - 30, 50);
		std::vector<Kmer> k1 = kmers_of(l1, s1);
		for (int i = n1, j = 0; j < n2; j++)
		{
			if (b[j].score < thr) continue;
			i64 left = b[j].diff - est, right = b[j].diff + l2;
			int e = chr_lookup(right); if (e < 0) continue;   // (UB in the reference)
			int cid = ix.ends[e].second;
			if (left < ix.G && left < ix.chr[cid].fwd_start - ix.chr[cid].len) left = ix.chr[cid].fwd_start - ix.chr[cid].len + 1;
			else if (right >= ix.G && left < ix.chr[cid].rev_start - ix.chr[cid].len) left = ix.chr[cid].rev_start - ix.chr[cid].len + 1;
			int slen = (int)(right - left); if (slen < l1) continue;
			std::string win = ref_slice(left, slen);
			std::vector<Kmer> k2 = kmers_of(slen, win.c_str());
			std::vector<Seg> sp = runs_to_pairs(10, common_kmers(slen, k1, k2));
			Cand c = rescue_in_window(left, sp);
			if (c.score > sc1) { mated = true; c.mate = j; b[j].mate = i++; a.push_back(c); }
		}
	}
	return mated;
}

static void settle_pair(ReadRes& r1, ReadRes& r2)   // CheckPairedFinalAlignments :429
{
	bool mated = r1.rep[r1.best].mate == r2.best;
	if (!g_pm.multihit && mated) return;
	if (!mated && r1.score > 0 && r2.score > 0)
	{
		int s = 0;
		for (int i = 0; i < r1.ncan; i++)
		{
			int j = r1.rep[i].mate;
			if (r1.rep[i].aln > 0 && j != -1 && r2.rep[j].aln > 0)
			{
				mated = true;
				if (s < r1.rep[i].aln + r2.rep[j].aln) { s = r1.rep[i].aln + r2.rep[j].aln; r1.best = i; r1.score = r1.rep[i].aln; r2.best = j; r2.score = r2.rep[j].aln; }
			}
		}
	}
	if (mated)
	{
		for (int i = 0; i < r1.ncan; i++)
		{
			int j = r1.rep[i].mate;
			if (r1.rep[i].aln != r1.score || (j != -1 && r2.rep[j].aln != r2.score)) { r1.rep[i].aln = 0; r1.rep[i].mate = -1; }
		}
	}
	else
	{
		for (int i = 0; i < r1.ncan; i++) { r1.rep[i].mate = -1; if (r1.rep[i].aln > 0 && r1.rep[i].aln != r1.score) r1.rep[i].aln = 0; }
		for (int j = 0; j < r2.ncan; j++) { r2.rep[j].mate = -1; if (r2.rep[j].aln > 0 && r2.rep[j].aln != r2.score) r2.rep[j].aln = 0; }
	}
}

static void flag_single(ReadRes& r)   // SetSingleAlignmentFlag :49
{
	if (r.score > r.sub) r.rep[r.best].flag = r.rep[r.best].fwd ? 0 : 0x10;
	else if (r.score > 0) { for (int i = 0; i < r.ncan; i++) if (r.rep[i].aln > 0) r.rep[i].flag = r.rep[i].fwd ? 0 : 0x10; }
	else r.rep[0].flag = 0x4;
}

static void flag_one_of_pair(ReadRes& me, ReadRes& other, int base)   // the two symmetric halves of SetPairedAlignmentFlag :96-156
{
	if (me.score > me.sub)
	{
		Report& a = me.rep[me.best]; a.flag = base | (a.fwd ? 0x20 : 0x10);
		if (a.mate != -1 && other.rep[a.mate].aln > 0) a.flag |= 0x2; else a.flag |= 0x8;
	}
	else if (me.score > 0)
	{
		for (int i = 0; i < me.ncan; i++)
		{
			Report& a = me.rep[i]; if (a.aln <= 0) continue;
			a.flag = base | (a.fwd ? 0x20 : 0x10);
			if (a.mate != -1 && other.rep[a.mate].aln > 0) a.flag |= 0x2; else a.flag |= 0x8;
		}
	}
	else
	{
		me.rep[0].flag = base | 0x4;
		if (other.score == 0) me.rep[0].flag |= 0x8; else me.rep[0].flag |= (other.rep[other.best].fwd ? 0x10 : 0x20);
	}
}

static void flag_pair(ReadRes& r1, ReadRes& r2)   // SetPairedAlignmentFlag :73
{
	if (r1.score > r1.sub && r2.score > r2.sub)
	{
		Report& a = r1.rep[r1.best]; Report& b = r2.rep[r2.best];
		a.flag = 0x41; b.flag = 0x81;
		if (r2.best == a.mate) { a.flag |= 0x2; b.flag |= 0x2; }
		a.flag |= a.fwd ? 0x20 : 0x10; b.flag |= b.fwd ? 0x20 : 0x10;
	}
	else { flag_one_of_pair(r1, r2, 0x41); flag_one_of_pair(r2, r1, 0x81); }
}

static void set_mapq(ReadRes& r)   // EvaluateMAPQ :160
{
	if (r.score == 0 || r.score == r.sub) { r.mapq = 0; return; }
	if (g_pm.pacbio)
	{
		float scale = 85.0 * (int)(ceil(r.rlen / 100 + 0.5));
		if (scale > 2000) scale = 2000;
		r.mapq = (int)(60 * (r.score / scale));
	}
	else if (r.sub == 0 || r.score - r.sub > 5) r.mapq = 60;
	else r.mapq = (int)(30 * (1 - (float)(r.score - r.sub) / r.score) * log(r.score) + 0.4999);
	if (r.mapq > 60) r.mapq = 60;
}

// ------------------------------------------------------------------------------------------------
// Per-read / per-pair drivers (Mapping.cpp:513-596)
// ------------------------------------------------------------------------------------------------
static void map_single(const char* seq, int rlen, ReadRes& r)
{
	std::vector<unsigned char> q; encode(seq, rlen, q); r = ReadRes(); r.rlen = rlen;
	std::vector<Cand> cv = g_pm.pacbio ? cands_pacbio(rlen, seeds_sensitive(rlen, q.data())) : cands_illumina(rlen, seeds_fast(rlen, q.data()));
	prune_candidates(cv);
	make_reports(true, seq, r, cv);
	flag_single(r); set_mapq(r);
}

struct PairExtra { bool paired, rescued; std::vector<Cand> c1, c2; int counted, absdist; };

static void map_pair(const char* s1, int l1, const char* s2, int l2, int est, ReadRes& r1, ReadRes& r2, PairExtra* ex)
{
	std::vector<unsigned char> q; r1 = ReadRes(); r2 = ReadRes(); r1.rlen = l1; r2.rlen = l2;
	encode(s1, l1, q); std::vector<Cand> a = cands_illumina(l1, seeds_fast(l1, q.data()));
	encode(s2, l2, q); std::vector<Cand> b = cands_illumina(l2, seeds_fast(l2, q.data()));   // (Mapping.cpp:550 encodes with l1; equal lengths assumed)
	bool paired = pair_candidates(est, a, b), rescued = false;
	if (!paired) { paired = rescue_pair(est, l1, s1, l2, s2, a, b); rescued = paired; }
	if (paired) keep_mated(a, b);
	prune_candidates(a); prune_candidates(b);
	if (ex) { ex->paired = paired; ex->rescued = rescued; ex->c1 = a; ex->c2 = b; }
	make_reports(true, s1, r1, a); make_reports(false, s2, r2, b);
	settle_pair(r1, r2);
	flag_pair(r1, r2);
	set_mapq(r1); set_mapq(r2);
	if (ex)
	{
		ex->counted = 0; ex->absdist = 0;
		if (r1.score > 0)
		{
			const Report& x = r1.rep[r1.best]; int j = x.mate;
			if (x.aln > 0 && j != -1 && r2.rep[j].aln > 0)
			{
				int dist = (int)(r2.rep[j].pos - x.pos + (x.fwd ? l2 : 0 - l1));
				ex->counted = 1; ex->absdist = abs(dist);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Text dumps (format in kart_oracle.h)
// ------------------------------------------------------------------------------------------------
static void put(std::string& s, const char* fmt, ...)
{
	va_list ap, ap2; va_start(ap, fmt); va_copy(ap2, ap);
	int n = vsnprintf(NULL, 0, fmt, ap); va_end(ap);
	std::vector<char> buf((size_t)n + 1); vsnprintf(buf.data(), buf.size(), fmt, ap2); va_end(ap2);
	s.append(buf.data(), n);
}
static long emit(const std::string& s, char* out, long cap) { if (out && (long)s.size() < cap) { memcpy(out, s.data(), s.size()); out[s.size()] = 0; } return (long)s.size(); }
static void dump_segs(std::string& s, const std::vector<Seg>& v) { for (size_t i = 0; i < v.size(); i++) put(s, "S %d %d %d %lld %d\n", v[i].rPos, v[i].rLen, v[i].gLen, (long long)v[i].gPos, v[i].simple ? 1 : 0); }
static void dump_cands(std::string& s, const std::vector<Cand>& v) { for (size_t i = 0; i < v.size(); i++) { put(s, "C %d %lld %d %d\n", v[i].score, (long long)v[i].diff, v[i].mate, (int)v[i].segs.size()); dump_segs(s, v[i].segs); } }
static void dump_read(std::string& s, const ReadRes& r)
{
	put(s, "R %d %d %d %d %d\n", r.score, r.sub, r.mapq, r.ncan, r.best);
	for (int i = 0; i < r.ncan; i++)
	{
		const Report& a = r.rep[i]; bool flagged = (r.score == 0 && i == 0) || (r.score > 0 && i == r.best);
		put(s, "A %d %d %d", i, a.aln, a.mate);
		if (flagged) put(s, " F%d", a.flag);
		if (a.aln > 0) put(s, " %d %d %lld %s", a.fwd ? 1 : 0, a.chr, (long long)a.pos, a.cigar.c_str());
		s += "\n";
	}
}

// ------------------------------------------------------------------------------------------------
// Whole-file driver: input (GetData.cpp), chunk loop + EstDistance (Mapping.cpp:488-622), SAM text (:177-315)
// ------------------------------------------------------------------------------------------------
struct Rd { std::string name, seq, qual; };

struct LineReader
{
	gzFile fp = NULL; std::string pending; bool has_pending = false;
	bool open(const char* fn) { fp = gzopen(fn, "rb"); if (fp) gzbuffer(fp, 1 << 20); return fp != NULL; }
	void close() { if (fp) gzclose(fp); fp = NULL; }
	bool line(std::string& s)   // one line including its '\n' (like getline / gzgets)
	{
		if (has_pending) { s = pending; has_pending = false; return true; }
		s.clear(); char buf[65536];
		while (gzgets(fp, buf, sizeof(buf)) != NULL) { s += buf; if (!s.empty() && s[s.size() - 1] == '\n') return true; }
		return !s.empty();
	}
	void unread(const std::string& s) { pending = s; has_pending = true; }
};

static bool next_entry(LineReader& in, bool fastq, Rd& r)   // GetNextEntry, GetData.cpp:51 (plain-file semantics)
{
	std::string ln; r = Rd();
	if (!in.line(ln)) return false;
	int len = (int)ln.size(), p1 = len - 1, p2 = len - 1;
	for (int i = 1; i < len; i++) if (ln[i] != '>' && ln[i] != '@') { p1 = i; break; }
	for (int i = 1; i < len; i++) if (ln[i] == ' ' || ln[i] == '/' || ln[i] == '\t') { p2 = i; break; }
	r.name = p2 > p1 ? ln.substr(p1, p2 - p1) : std::string();
	if (fastq)
	{
		std::string sq, plus, ql;
		if (!in.line(sq)) return false;
		in.line(plus); in.line(ql);
		int rl = (int)sq.size() - 1; if (rl <= 0) return false;
		r.seq = sq.substr(0, rl); ql.resize(sq.size(), '\0'); r.qual = std::string(ql.substr(0, rl).c_str());
	}
	else
	{
		while (in.line(ln))
		{
			if (ln[0] == '>') { in.unread(ln); break; }
			r.seq += ln.substr(0, ln.size() - 1);
		}
		if (r.seq.empty()) return false;
	}
	return true;
}

static char comp(char c) { switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; } }
static std::string revcomp(const std::string& s) { std::string o(s.size(), 'N'); for (size_t i = 0; i < s.size(); i++) o[s.size() - 1 - i] = comp(s[i]); return o; }

static void sam_unmapped(std::string& o, const Rd& rd, const ReadRes& r, bool fastq)
{
	char b[64]; o += rd.name; snprintf(b, sizeof(b), "\t%d\t*\t0\t0\t*\t*\t0\t0\t", r.rep[0].flag); o += b; o += rd.seq; o += '\t'; o += fastq ? rd.qual : "*"; o += "\tAS:i:0\tXS:i:0\n";
}

static void sam_mapped(std::string& o, const Rd& rd, const ReadRes& r, const Report& a, bool stored_fwd, bool fastq, bool has_mate, i64 mate_pos, int tlen)
{
	// stored_fwd: orientation of rd.seq as kept in memory (mate 2 is held reverse-complemented, GetData.cpp:125)
	char b[128]; bool as_is = (a.fwd == stored_fwd);
	o += rd.name; snprintf(b, sizeof(b), "\t%d\t", a.flag); o += b; o += g_ix.chr[a.chr].name;
	snprintf(b, sizeof(b), "\t%lld\t%d\t", (long long)a.pos, r.mapq); o += b; o += a.cigar;
	if (has_mate) { snprintf(b, sizeof(b), "\t=\t%lld\t%d\t", (long long)mate_pos, tlen); o += b; } else o += "\t*\t0\t0\t";
	o += as_is ? rd.seq : revcomp(rd.seq); o += '\t';
	if (!fastq) o += "*"; else if (as_is) o += rd.qual; else o.append(rd.qual.rbegin(), rd.qual.rend());
	snprintf(b, sizeof(b), "\tNM:i:%d\tAS:i:%d\tXS:i:%d\n", r.rlen - r.score, r.score, r.sub); o += b;
}

static void sam_pair(std::string& o, const Rd& d1, const Rd& d2, const ReadRes& r1, const ReadRes& r2, bool fastq, i64& n_paired, i64& sum_dist)   // OutputPairedAlignments :177
{
	if (r1.score == 0) sam_unmapped(o, d1, r1, fastq);
	else for (int i = r1.best; i < r1.ncan; i++)
	{
		const Report& a = r1.rep[i];
		if (a.aln > 0)
		{
			int j = a.mate;
			if (j != -1 && r2.rep[j].aln > 0)
			{
				int dist = (int)(r2.rep[j].pos - a.pos + (a.fwd ? r2.rlen : 0 - r1.rlen));
				if (i == r1.best) { n_paired += 2; if (abs(dist) < 10000) sum_dist += abs(dist); }
				sam_mapped(o, d1, r1, a, true, fastq, true, r2.rep[j].pos, dist);
			}
			else sam_mapped(o, d1, r1, a, true, fastq, false, 0, 0);
		}
		if (!g_pm.multihit) break;
	}
	if (r2.score == 0) sam_unmapped(o, d2, r2, fastq);
	else for (int j = r2.best; j < r2.ncan; j++)
	{
		const Report& b = r2.rep[j];
		if (b.aln > 0)
		{
			int i = b.mate;
			if (i != -1 && r1.rep[i].aln > 0)
			{
				int dist = 0 - (int)(b.pos - r1.rep[i].pos + (r1.rep[i].fwd ? r2.rlen : 0 - r1.rlen));
				sam_mapped(o, d2, r2, b, false, fastq, true, r1.rep[i].pos, dist);
			}
			else sam_mapped(o, d2, r2, b, false, fastq, false, 0, 0);
		}
		if (!g_pm.multihit) break;
	}
}

static void sam_single(std::string& o, const Rd& d, const ReadRes& r, bool fastq)   // OutputSingledAlignments :272
{
	if (r.score == 0) { sam_unmapped(o, d, r, fastq); return; }
	for (int i = r.best; i < r.ncan; i++)
		if (r.rep[i].aln == r.score) { sam_mapped(o, d, r, r.rep[i], true, fastq, false, 0, 0); if (!g_pm.multihit) break; }
}

extern "C" {

int kor_load(const char* prefix, int pacbio, int max_gaps, int multihit) { g_pm = Params(); g_pm.pacbio = pacbio != 0; g_pm.max_gaps = max_gaps; g_pm.multihit = multihit != 0; return load_index(prefix); }
void kor_set_mode(int pacbio, int max_gaps, int multihit) { g_pm.pacbio = pacbio != 0; g_pm.max_gaps = max_gaps; g_pm.multihit = multihit != 0; }
int kor_min_seed_len(void) { return g_pm.min_seed; }
long long kor_genome_size(void) { return g_ix.G; }
void kor_occ4(unsigned long long k, unsigned long long* cnt) { u64 c[4]; occ4(k, c); for (int i = 0; i < 4; i++) cnt[i] = c[i]; }
unsigned long long kor_sa(unsigned long long k) { return sa_locate(k); }
void kor_counters(unsigned long long* c, int reset) { for (int i = 0; i < 7; i++) { c[i] = g_cnt[i]; if (reset) g_cnt[i] = 0; } }

void kor_bwt_search(const uint8_t* codes, int start, int stop, int* len, int* freq, unsigned long long* locs, unsigned long long* x0, unsigned long long* x2)
{
	SearchHit h = fm_search(codes, start, stop);
	*len = h.len; *freq = h.freq; if (x0) *x0 = h.x0; if (x2) *x2 = h.x2;
	for (int i = 0; i < h.freq; i++) locs[i] = h.loc[i];
}

long kor_seeds(const char* seq, int rlen, int sensitive, char* out, long cap)
{
	std::vector<unsigned char> q; encode(seq, rlen, q); std::string s;
	dump_segs(s, sensitive ? seeds_sensitive(rlen, q.data()) : seeds_fast(rlen, q.data()));
	return emit(s, out, cap);
}

long kor_candidates(const char* seq, int rlen, char* out, long cap)
{
	std::vector<unsigned char> q; encode(seq, rlen, q); std::string s;
	dump_cands(s, g_pm.pacbio ? cands_pacbio(rlen, seeds_sensitive(rlen, q.data())) : cands_illumina(rlen, seeds_fast(rlen, q.data())));
	return emit(s, out, cap);
}

long kor_nw(const char* s1, int m, const char* s2, int n, char* o1, char* o2)
{
	std::string a(s1, m), b(s2, n); nw_align(a, b);
	memcpy(o1, a.c_str(), a.size() + 1); memcpy(o2, b.c_str(), b.size() + 1);
	return (long)a.size();
}

long kor_fragment_pairs(int max_dist, const char* f1, int len1, const char* f2, int len2, int do_normal, char* out, long cap)
{
	std::string a(f1, len1), b(f2, len2), s;
	std::vector<Seg> v = fragment_pairs(max_dist, len1, a.c_str(), len2, b.c_str());
	if (do_normal && v.size() > 0) fill_pairs(len1, len2, v);
	dump_segs(s, v);
	return emit(s, out, cap);
}

long kor_normal_pairs(int rlen, int glen, int n, const int* rpos, const int* rl, const long long* gpos, char* out, long cap)
{
	std::vector<Seg> v(n); std::string s;
	for (int i = 0; i < n; i++) { v[i].simple = true; v[i].rPos = rpos[i]; v[i].rLen = v[i].gLen = rl[i]; v[i].gPos = gpos[i]; v[i].diff = gpos[i] - rpos[i]; }
	fill_pairs(rlen, glen, v); dump_segs(s, v);
	return emit(s, out, cap);
}

long kor_process_pair(int kind, const char* seq, int rpos, int rlen, long long gpos, int glen, char* out, long cap)
{
	Seg sp; sp.simple = false; sp.rPos = rpos; sp.rLen = rlen; sp.gPos = gpos; sp.gLen = glen; sp.diff = gpos - rpos;
	Cigar cg; std::string s;
	int score = kind == 0 ? do_middle(seq, sp, cg) : (kind == 1 ? do_head(seq, sp, cg) : do_tail(seq, sp, cg));
	put(s, "P %d %d %d %lld %d\n", score, sp.rPos, sp.rLen, (long long)sp.gPos, sp.gLen);
	for (size_t i = 0; i < cg.size(); i++) put(s, "O %d %c\n", cg[i].first, cg[i].second);
	return emit(s, out, cap);
}

long kor_map_single(const char* seq, int rlen, char* out, long cap)
{
	std::string sq(seq, rlen), s; ReadRes r; map_single(sq.c_str(), rlen, r); dump_read(s, r);
	return emit(s, out, cap);
}

long kor_map_pair(const char* seq1, int l1, const char* seq2rc, int l2, int est, int stage_dump, char* out, long cap)
{
	std::string a(seq1, l1), b(seq2rc, l2), s; ReadRes r1, r2; PairExtra ex;
	map_pair(a.c_str(), l1, b.c_str(), l2, est, r1, r2, &ex);
	if (stage_dump) { put(s, "X %d %d\n", ex.paired ? 1 : 0, ex.rescued ? 1 : 0); s += "V1\n"; dump_cands(s, ex.c1); s += "V2\n"; dump_cands(s, ex.c2); }
	dump_read(s, r1); dump_read(s, r2);
	put(s, "P %d %d\n", ex.counted, ex.absdist);
	return emit(s, out, cap);
}

long kor_map_files(const char* f1, const char* f2, int paired, const char* out_sam, int threads)
{
	if (!g_ix.loaded) return -1;
	LineReader in1, in2; if (!in1.open(f1)) return -2;
	bool sep = f2 != NULL && f2[0] != 0; if (sep && !in2.open(f2)) { in1.close(); return -2; }
	bool pe = paired != 0 || sep;
	bool fastq; { gzFile t = gzopen(f1, "rb"); char c = 0; gzread(t, &c, 1); gzclose(t); fastq = c == '@'; }   // CheckReadFormat, GetData.cpp:8
	FILE* out = fopen(out_sam, "w"); if (!out) { in1.close(); in2.close(); return -3; }
	fprintf(out, "@PG\tID:kart\tPN:Kart\tVN:2.5.6\n");
	for (size_t i = 0; i < g_ix.chr.size(); i++) fprintf(out, "@SQ\tSN:%s\tLN:%lld\n", g_ix.chr[i].name.c_str(), (long long)g_ix.chr[i].len);
	const int chunk_cap = g_pm.pacbio ? 10 : 4000;
	i64 total = 0, n_paired = 0, sum_dist = 0; if (threads < 1) threads = 1;
	while (true)
	{
		std::vector<Rd> rd;   // GetNextChunk, GetData.cpp:109
		while (true)
		{
			Rd a, b;
			if (!next_entry(in1, fastq, a)) break;
			rd.push_back(a);
			if (!next_entry(sep ? in2 : in1, fastq, b)) break;
			if (pe) { b.seq = revcomp(b.seq); std::reverse(b.qual.begin(), b.qual.end()); }
			rd.push_back(b);
			if ((int)rd.size() == chunk_cap) break;
		}
		int n = (int)rd.size(); if (n == 0) break;
		bool as_pairs = !g_pm.pacbio && pe && n % 2 == 0;
		int est = 1500;
		if (n_paired >= 1000) { est = (int)(sum_dist / (n_paired >> 2)); est += est >> 1; }   // Mapping.cpp:533-540
		std::vector<ReadRes> rr(n);
		std::atomic<int> next(0);
		auto work = [&]() {
			if (as_pairs) { int i; while ((i = next.fetch_add(2)) < n) map_pair(rd[i].seq.c_str(), (int)rd[i].seq.size(), rd[i + 1].seq.c_str(), (int)rd[i + 1].seq.size(), est, rr[i], rr[i + 1], NULL); }
			else { int i; while ((i = next.fetch_add(1)) < n) map_single(rd[i].seq.c_str(), (int)rd[i].seq.size(), rr[i]); }
		};
		if (threads == 1) work();
		else { std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(work); for (auto& t : th) t.join(); }
		std::string o;
		if (as_pairs) for (int i = 0; i < n; i += 2) sam_pair(o, rd[i], rd[i + 1], rr[i], rr[i + 1], fastq, n_paired, sum_dist);
		else for (int i = 0; i < n; i++) sam_single(o, rd[i], rr[i], fastq);
		fwrite(o.data(), 1, o.size(), out);
		total += n;
	}
	fclose(out); in1.close(); in2.close();
	return (long)total;
}

} // extern "C"
