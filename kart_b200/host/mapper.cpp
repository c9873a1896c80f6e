// Batch driver: the host half of ReadMapping()/Mapping() (reference src/Mapping.cpp:488-742).
//
// The reference maps 4000-read chunks; chunk c uses EstDistance derived from the insert-size statistic of chunks 0..c-1
// (:533-540, exact for `-t 1`). Here many chunks are mapped per GPU launch with a predicted EstDistance, every pair reports
// the interval of EstDistance values for which its result cannot change (kb_pair_stat_t::est_lo/est_hi), and this file
// replays the per-chunk recurrence in input order, re-mapping only the pairs whose interval excludes the true value, until
// the batch is self-consistent. The output is therefore identical to `kart -t 1`.
#include "kart_host.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <time.h>

struct BatchResult { std::vector<kb_aln_t> aln; std::vector<kb_pair_stat_t> pairs; std::vector<uint32_t> cigar; std::vector<int32_t> est_used; };

static int map_batch(kb_ctx_t* ctx, const uint8_t* seq, const uint64_t* off, int n, const int32_t* est, BatchResult& out)
{
	kb_reads_t in; in.n_reads = n; in.seq = seq; in.seq_off = off;
	out.aln.resize(n); out.pairs.resize(n / 2 + 1);
	if (out.cigar.size() < (size_t)n * 4 + 1024) out.cigar.resize((size_t)n * 4 + 1024);
	kb_results_t res; res.aln = out.aln.data(); res.pairs = out.pairs.data(); res.cigar = out.cigar.data(); res.cap_cigar = (uint32_t)out.cigar.size(); res.n_cigar = 0;
	int rc = kb_map_chunk(ctx, &in, est, &res);
	if (rc == KB_ECAPACITY)
	{
		out.cigar.resize((size_t)res.n_cigar + 1024); res.cigar = out.cigar.data(); res.cap_cigar = (uint32_t)out.cigar.size();
		rc = kb_fetch_results(ctx, &res);
	}
	if (rc == KB_OK) out.cigar.resize(res.n_cigar);
	return rc;
}

struct PairState { long long iPaired = 0, iDistance = 0; };
static inline int est_of(const PairState& s) { if (s.iPaired >= 1000) { int e = (int)(s.iDistance / (s.iPaired >> 2)); return e + (e >> 1); } return 1500; }

// Replays the chunk recurrence over one mapped batch; re-maps the pairs whose result depends on the difference between the
// predicted and the true EstDistance. Returns 0 or a kb error code.
static int settle_est(kb_ctx_t* ctx, const ReadBatch& b, BatchResult& br, PairState& st, int chunk_reads, long long* remapped)
{
	int n = b.n(), np = n / 2;
	while (true)
	{
		PairState s = st; std::vector<int> viol; std::vector<int32_t> viol_est;
		for (int c0 = 0; c0 < n; c0 += chunk_reads)
		{
			int c1 = std::min(n, c0 + chunk_reads), est = est_of(s);
			for (int p = c0 / 2; p < c1 / 2; p++)
			{
				const kb_pair_stat_t& ps = br.pairs[p];
				if (br.est_used[p] != est && (est < ps.est_lo || est > ps.est_hi)) { viol.push_back(p); viol_est.push_back(est); }
				if (ps.counted) { s.iPaired += 2; if (ps.absdist < 10000) s.iDistance += ps.absdist; }
			}
		}
		if (viol.empty()) { st = s; return KB_OK; }
		*remapped += (long long)viol.size();
		// gather the affected pairs into a small batch and map them again with their true EstDistance
		std::vector<uint8_t> seq; std::vector<uint64_t> off(1, 0);
		for (int p : viol) for (int r = 2 * p; r < 2 * p + 2; r++) { seq.insert(seq.end(), b.seq.begin() + b.seq_off[r], b.seq.begin() + b.seq_off[r + 1]); off.push_back(seq.size()); }
		BatchResult fix;
		int rc = map_batch(ctx, seq.data(), off.data(), (int)off.size() - 1, viol_est.data(), fix); if (rc) return rc;
		for (size_t k = 0; k < viol.size(); k++)
		{
			int p = viol[k];
			for (int h = 0; h < 2; h++)
			{
				kb_aln_t a = fix.aln[2 * k + h];
				uint32_t at = (uint32_t)br.cigar.size();
				br.cigar.insert(br.cigar.end(), fix.cigar.begin() + a.cig_off, fix.cigar.begin() + a.cig_off + a.cig_len);
				a.cig_off = at; br.aln[2 * p + h] = a;
			}
			br.pairs[p] = fix.pairs[k]; br.est_used[p] = viol_est[k];
		}
		(void)np;
	}
}

int run_mapping(const RunOptions& opt, const HostIndex& idx)
{
	kb_ctx_t* ctx = nullptr;
	int rc = kb_init(0, &ctx);
	if (rc) { fprintf(stderr, "Error! kart_b200 needs a CUDA device: %s\n", kb_strerror(rc)); return 1; }
	kb_index_host_t hi; idx.describe(&hi);
	if ((rc = kb_upload_index(ctx, &hi, opt.expand_sa ? 1 : 0)) != 0) { fprintf(stderr, "Error! index upload failed: %s (%s)\n", kb_strerror(rc), kb_last_error(ctx)); kb_destroy(ctx); return 1; }

	FILE* out = nullptr; BamWriter bam; const bool to_bam = opt.out_format == 1 && !opt.debug;
	std::vector<int32_t> name2id(idx.chr_name.size(), -1);   // bam_name2id: a repeated @SQ name keeps its first id (htslib sam.c:725-733)
	if (!opt.debug)
	{
		std::string h; sam_header(h, idx);
		if (to_bam)
		{
			if (!bam.open(opt.out_name.c_str(), opt.threads)) { fprintf(stderr, "Error! Cannot open file [%s]\n", opt.out_name.c_str()); kb_destroy(ctx); exit(1); }
			std::vector<std::string> un; std::vector<int64_t> ul;
			for (size_t i = 0; i < idx.chr_name.size(); i++)
			{
				size_t k = 0; while (k < un.size() && un[k] != idx.chr_name[i]) k++;
				if (k == un.size()) { un.push_back(idx.chr_name[i]); ul.push_back(idx.chr_len[i]); }
				name2id[i] = (int32_t)k;
			}
			bam_tables_init(); bam.write_header(h, un, ul);
		}
		else
		{
			out = fopen(opt.out_name.c_str(), "w");
			if (!out) { fprintf(stderr, "Error! Cannot open file [%s]\n", opt.out_name.c_str()); kb_destroy(ctx); exit(1); }
			fwrite(h.data(), 1, h.size(), out);
		}
	}
	if (opt.silent) fprintf(stdout, "Start read mapping...\n");
	time_t t0 = time(NULL);
	long long total = 0, unmapped = 0, unique = 0, remapped = 0; PairState st; bool pair_end = opt.pair_flag;
	const int chunk_reads = opt.pacbio ? 10 : 4000;
	int fmt_threads = std::max(1, opt.threads);

	for (size_t lib = 0; lib < opt.files1.size(); lib++)
	{
		ReadSource src; bool sep = opt.files1.size() == opt.files2.size();
		if (sep) pair_end = true;
		if (sep)
		{
			// both files must have the same format (Mapping.cpp:700-709)
			ReadSource a, b2; bool ok1 = a.open(opt.files1[lib].c_str(), nullptr), ok2 = b2.open(opt.files2[lib].c_str(), nullptr);
			bool same = ok1 && ok2 && a.fastq == b2.fastq; a.close(); b2.close();
			if (ok1 && ok2 && !same) { fprintf(stdout, "Error! %s and %s are with different format...\n", opt.files1[lib].c_str(), opt.files2[lib].c_str()); continue; }
			if (!ok1 || !ok2) continue;
		}
		if (!src.open(opt.files1[lib].c_str(), sep ? opt.files2[lib].c_str() : nullptr)) continue;
		bool fastq = src.fastq;
		kb_params_t pm; pm.min_seed_len = 0; pm.max_gaps = opt.max_gaps; pm.max_insert = 1500; pm.pacbio = opt.pacbio; pm.multihit = opt.multihit;

		// pipeline: reader thread fills batch k+1 while the GPU maps batch k
		ReadBatch cur, nxt; cur.clear(); nxt.clear();
		int batch_reads = std::max(chunk_reads, opt.batch_reads / chunk_reads * chunk_reads);
		int got = src.fill(cur, batch_reads, pair_end);
		while (got > 0)
		{
			std::thread reader([&]() { nxt.clear(); src.fill(nxt, batch_reads, pair_end); });
			if (!opt.silent) { fprintf(stdout, "\r%lld %s reads have been processed in %ld seconds...", total, pair_end ? "paired-end" : "singled-end", (long)(time(NULL) - t0)); fflush(stdout); }
			int n = cur.n();
			// a batch with an odd number of reads can only be the last one: its final read goes through the single-end branch (Mapping.cpp:531,598)
			// (the reference sends a whole chunk through the single-end branch when its read count is odd: the final short chunk)
			int n_pe = (!opt.pacbio && pair_end) ? ((n & 1) ? (n / chunk_reads) * chunk_reads : n) : 0;
			BatchResult br, tail;
			if (n_pe > 0)
			{
				pm.paired = 1; kb_set_params(ctx, &pm);
				br.est_used.assign(n_pe / 2, est_of(st));
				rc = map_batch(ctx, cur.seq.data(), cur.seq_off.data(), n_pe, br.est_used.data(), br);
				if (!rc) rc = settle_est(ctx, cur, br, st, chunk_reads, &remapped);
			}
			if (!rc && n > n_pe)
			{
				pm.paired = 0; kb_set_params(ctx, &pm);
				std::vector<uint64_t> off(cur.seq_off.begin() + n_pe, cur.seq_off.end());
				uint64_t base = off[0]; for (auto& o : off) o -= base;
				rc = map_batch(ctx, cur.seq.data() + base, off.data(), n - n_pe, nullptr, tail);
			}
			if (rc) { fprintf(stderr, "\nError! GPU mapping failed: %s (%s)\n", kb_strerror(rc), kb_last_error(ctx)); reader.join(); break; }
			// format SAM in parallel slices, write in input order
			std::vector<std::string> parts(fmt_threads); std::vector<std::thread> th; std::vector<std::vector<uint32_t>> rec_ends(fmt_threads);
			std::vector<long long> um(fmt_threads, 0), uq(fmt_threads, 0);
			for (int t = 0; t < fmt_threads; t++) th.emplace_back([&, t]() {
				int lo = (int)((long long)n * t / fmt_threads), hi = (int)((long long)n * (t + 1) / fmt_threads);
				if (n_pe) { lo &= ~1; if (t + 1 < fmt_threads) hi &= ~1; }
				std::string& o = parts[t]; o.reserve((size_t)(hi - lo) * 400);
				for (int r = lo; r < hi; r++)
				{
					bool in_pe = r < n_pe;
					const kb_aln_t& a = in_pe ? br.aln[r] : tail.aln[r - n_pe];
					const uint32_t* cg = in_pe ? br.cigar.data() : tail.cigar.data();
					if (a.score == 0) um[t]++; else if (a.mapq == 60) uq[t]++;
					if (to_bam) bam_read_record(o, rec_ends[t], name2id, cur, r, !(in_pe && (r & 1)), a, cg, fastq);
					else sam_read_line(o, idx, cur, r, !(in_pe && (r & 1)), a, cg, fastq);   // mate 2 of a mapped pair is held reverse-complemented
				}
			});
			for (auto& t : th) t.join();
			if (out) for (auto& p : parts) fwrite(p.data(), 1, p.size(), out);
			if (to_bam) for (int t = 0; t < fmt_threads; t++) bam.append(parts[t], rec_ends[t], true);
			for (int t = 0; t < fmt_threads; t++) { unmapped += um[t]; unique += uq[t]; }
			total += n;
			reader.join();
			std::swap(cur, nxt); got = cur.n();
		}
		src.close();
	}
	fprintf(stdout, "\rAll the %lld %s reads have been processed in %lld seconds.\n", total, pair_end ? "paired-end" : "single-end", (long long)(time(NULL) - t0));
	if (out) fclose(out);
	if (to_bam) bam.close();
	if (total > 0)
	{
		if (pair_end) fprintf(stdout, "\t# of total mapped sequences = %lld (sensitivity = %.2f%%)\n\t# of paired sequences = %lld (%.2f%%), average insert size = %d\n", total - unmapped, (int)(10000 * (1.0 * (total - unmapped) / total) + 0.5) / 100.0, st.iPaired, (int)(10000 * (1.0 * st.iPaired / total) + 0.5) / 100.0, (st.iPaired > 1 ? (int)(st.iDistance / (st.iPaired >> 1)) : 0));
		else fprintf(stdout, "\t# of total mapped sequences = %lld (sensitivity = %.2f%%)\n", total - unmapped, (int)(10000 * (1.0 * (total - unmapped) / total) + 0.5) / 100.0);
		fprintf(stdout, "Alignment output: %s\n", opt.out_name.c_str());
	}
	(void)unique; (void)remapped;
	kb_destroy(ctx);
	return rc ? 1 : 0;
}
