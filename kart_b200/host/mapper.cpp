// Batch driver: the host half of ReadMapping()/Mapping() (reference src/Mapping.cpp:488-742).
//
// The reference maps 4000-read chunks; chunk c uses EstDistance derived from the insert-size statistic of chunks 0..c-1
// (:533-540, exact for `-t 1`). Here many chunks are mapped per GPU launch with a predicted EstDistance, every pair reports
// the interval of EstDistance values for which its result cannot change (kb_pair_stat_t::est_lo/est_hi), and this file
// replays the per-chunk recurrence in input order, re-mapping only the pairs whose interval excludes the true value, until
// the batch is self-consistent. The output is therefore identical to `kart -t 1`.
//
// Several GPUs (the reference's worker pool, pthread_create x iThreadNum over one chunk source, Mapping.cpp:716-717,504-512): one
// host thread and one kb_ctx_t per device take batches from the same queue in input order. Mapping a batch needs nothing from the
// other devices; the recurrence does, so a worker settles its batch only when every earlier batch has been settled (a turnstile
// on the batch number) and hands it to the writer in that order. The first device gets the index from the host, the others copy
// it from the first one's HBM (kb_clone_index: NVLink between peers). Devices beyond the first are only brought up when the
// input is long enough to need them (a CUDA context costs most of a second).
#include "kart_host.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <deque>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>

struct BatchResult
{
	HBuf<kb_aln_t> aln{true}; HBuf<kb_pair_stat_t> pairs{true}; HBuf<uint32_t> cigar{true}; std::vector<int32_t> est_used;
	std::vector<kb_extra_t> extra;   // -m: further lines, sorted by (read, rank)
};
static bool g_multihit = false;
// KART_B200_TRACE=1: wall time of every pipeline stage per batch on stderr (reader fill, GPU map, EstDistance settle, format, write)
static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
static const bool g_trace = getenv("KART_B200_TRACE") != nullptr;
static double g_t0 = 0;

// Packs the reads on the host threads (2 bits per base + the list of characters that are no upper-case bases) and maps them
// through kb_map_chunk_packed: a third of the bytes of the text over PCIe. `pk` holds the page-locked staging of one worker.
struct PackBuf { HBuf<uint64_t> code{true}; HBuf<uint64_t> exc{true}; };
static int g_pack_threads = 4;
static int map_batch(kb_ctx_t* ctx, const uint8_t* seq, const uint64_t* off, int n, const int32_t* est, BatchResult& out, PackBuf& pk)
{
	const double t_in = g_trace ? now_s() : 0;
	kb_reads_t in; in.n_reads = n; in.seq = seq; in.seq_off = off;
	out.aln.resize(n); out.pairs.resize(n / 2 + 1);
	if (out.cigar.cap < (size_t)n * 4 + 1024) { out.cigar.clear(); out.cigar.reserve((size_t)n * 4 + 1024); }
	kb_results_t res; res.aln = out.aln.data(); res.pairs = out.pairs.data(); res.cigar = out.cigar.data(); res.cap_cigar = (uint32_t)out.cigar.cap; res.n_cigar = 0;
	kb_reads_packed_t pr;
	pk.code.resize((size_t)kb_packed_words(&in));
	if (pk.exc.cap < 4096) pk.exc.reserve(4096);
	const double t_a = g_trace ? now_s() : 0;
	int rc = kb_pack_reads(&in, pk.code.data(), pk.exc.data(), pk.exc.cap, g_pack_threads, &pr);
	if (rc == KB_ECAPACITY) { pk.exc.clear(); pk.exc.reserve((size_t)pr.n_exc + 4096); rc = kb_pack_reads(&in, pk.code.data(), pk.exc.data(), pk.exc.cap, g_pack_threads, &pr); }
	const double t_b = g_trace ? now_s() : 0;
	if (rc == KB_OK) rc = kb_map_chunk_packed(ctx, &pr, est, &res);
	if (g_trace && n > 100000) fprintf(stderr, "[kart trace]   map_batch %8d reads: buffers %.1f ms, pack %.1f ms, kb_map_chunk_packed %.1f ms\n", n, (t_a - t_in) * 1e3, (t_b - t_a) * 1e3, (now_s() - t_b) * 1e3);
	if (rc == KB_ECAPACITY)
	{
		out.cigar.clear(); out.cigar.reserve((size_t)res.n_cigar + 1024); res.cigar = out.cigar.data(); res.cap_cigar = (uint32_t)out.cigar.cap;
		rc = kb_fetch_results(ctx, &res);
	}
	if (rc == KB_OK) out.cigar.n = res.n_cigar;
	out.extra.clear();
	if (rc == KB_OK && g_multihit)
	{
		uint32_t ne = 0; rc = kb_fetch_extra(ctx, nullptr, 0, &ne);
		if (rc == KB_ECAPACITY) { out.extra.resize(ne); rc = kb_fetch_extra(ctx, out.extra.data(), ne, &ne); }
	}
	return rc;
}


struct PairState { long long iPaired = 0, iDistance = 0; };
static inline int est_of(const PairState& s) { if (s.iPaired >= 1000) { int e = (int)(s.iDistance / (s.iPaired >> 2)); return e + (e >> 1); } return 1500; }

// Replays the chunk recurrence over one mapped batch; re-maps the pairs whose result depends on the difference between the
// predicted and the true EstDistance. Returns 0 or a kb error code.
static int settle_est(kb_ctx_t* ctx, const ReadBatch& b, BatchResult& br, PairState& st, int chunk_reads, int n, long long* remapped, PackBuf& pk)
{
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
		for (int p : viol) for (int r = 2 * p; r < 2 * p + 2; r++) { seq.insert(seq.end(), b.seq.data() + b.seq_off[r], b.seq.data() + b.seq_off[r + 1]); off.push_back(seq.size()); }
		BatchResult fix;
		int rc = map_batch(ctx, seq.data(), off.data(), (int)off.size() - 1, viol_est.data(), fix, pk); if (rc) return rc;
		const uint32_t base = (uint32_t)br.cigar.size();   // the re-mapped pairs' cigar elements go behind the batch's, offsets shift by base
		br.cigar.append(fix.cigar.data(), fix.cigar.size());
		for (size_t k = 0; k < viol.size(); k++)
		{
			int p = viol[k];
			for (int h = 0; h < 2; h++) { kb_aln_t a = fix.aln[2 * k + h]; a.cig_off += base; br.aln[2 * p + h] = a; }
			br.pairs[p] = fix.pairs[k]; br.est_used[p] = viol_est[k];
		}
		if (g_multihit)
		{
			std::vector<char> redo((size_t)n / 2 + 1, 0); for (int p : viol) redo[p] = 1;
			std::vector<kb_extra_t> keep; keep.reserve(br.extra.size() + fix.extra.size());
			for (const kb_extra_t& e : br.extra) if (!redo[e.read >> 1]) keep.push_back(e);
			for (kb_extra_t e : fix.extra) { e.read = (uint32_t)(2 * viol[e.read >> 1]) + (e.read & 1); e.aln.cig_off += base; keep.push_back(e); }
			std::sort(keep.begin(), keep.end(), [](const kb_extra_t& a, const kb_extra_t& b) { return a.read != b.read ? a.read < b.read : a.rank < b.rank; });
			br.extra.swap(keep);
		}
	}
}

// One batch travelling through the three stages (reader thread -> GPU on the main thread -> writer thread).
struct Job
{
	ReadBatch rb; BatchResult br, tail; int n_pe = 0; bool last_of_run = false; long long seq_no = 0;
};
template <class T> class Channel   // small blocking queue
{
public:
	void put(T v) { { std::lock_guard<std::mutex> g(m); q.push_back(v); } cv.notify_one(); }
	T take() { std::unique_lock<std::mutex> g(m); cv.wait(g, [&]() { return !q.empty(); }); T v = q.front(); q.pop_front(); return v; }
private:
	std::mutex m; std::condition_variable cv; std::deque<T> q;
};

// Output file: header and batches in input order; a batch's slices are written with parallel pwrite when the file is seekable.
struct SamSink
{
	int fd = -1; bool seekable = false; off_t at = 0;
	bool open(const char* path)
	{
		fd = ::open(path, O_CREAT | O_TRUNC | O_WRONLY, 0666); if (fd < 0) return false;
		struct stat st; seekable = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
		return true;
	}
	static void write_all(int fd, const char* p, size_t n) { while (n) { ssize_t w = ::write(fd, p, n); if (w <= 0) { fprintf(stderr, "Error! write failed\n"); exit(1); } p += w; n -= (size_t)w; } }
	static void pwrite_all(int fd, const char* p, size_t n, off_t o) { while (n) { ssize_t w = ::pwrite(fd, p, n > (64u << 20) ? (64u << 20) : n, o); if (w <= 0) { fprintf(stderr, "Error! write failed\n"); exit(1); } p += w; n -= (size_t)w; o += w; } }
	void write(const char* p, size_t n) { write_all(fd, p, n); at += (off_t)n; if (seekable) lseek(fd, at, SEEK_SET); }
	void write_parts(std::vector<HBuf<char>>& parts)
	{
		if (!seekable) { for (auto& p : parts) write_all(fd, p.data(), p.size()); return; }
		std::vector<off_t> o(parts.size() + 1, at); for (size_t i = 0; i < parts.size(); i++) o[i + 1] = o[i] + (off_t)parts[i].size();
		std::vector<std::thread> th;
		for (size_t i = 1; i < parts.size(); i++) th.emplace_back([&, i]() { pwrite_all(fd, parts[i].data(), parts[i].size(), o[i]); });
		if (!parts.empty()) pwrite_all(fd, parts[0].data(), parts[0].size(), o[0]);
		for (auto& t : th) t.join();
		at = o[parts.size()]; lseek(fd, at, SEEK_SET);
	}
	void close() { if (fd >= 0) ::close(fd); fd = -1; }
};

int run_mapping(const RunOptions& opt, const HostIndex& idx)
{
	const int chunk_reads = opt.pacbio ? 10 : 4000;
	const int io_threads = std::max(1, opt.threads);
	const int batch_reads = std::max(chunk_reads, std::min(opt.batch_reads, opt.pacbio ? (1 << 16) : (1 << 30)) / chunk_reads * chunk_reads);   // long reads: bound the bases per batch

	// stage 1 starts before the device is ready: the first batch is parsed while the index is uploaded
	const int n_dev_max = opt.n_gpus > 0 ? opt.n_gpus : std::max(1, kb_device_count());   // default: every visible device
	const int n_jobs = 2 * n_dev_max + 2;
	g_t0 = now_s();
	std::vector<Job> jobs(n_jobs);
	Channel<Job*> free_q, ready_q, done_q;
	for (auto& j : jobs) free_q.put(&j);
	bool pair_end_final = opt.pair_flag;
	std::thread reader([&]() {
		bool pair_end = opt.pair_flag;
		for (size_t lib = 0; lib < opt.files1.size(); lib++)
		{
			ReadSource src; src.threads = io_threads; bool sep = opt.files1.size() == opt.files2.size();
			if (sep) pair_end = true;
			if (sep)
			{
				// both files must have the same format (Mapping.cpp:700-709)
				ReadSource a, b2; bool ok1 = a.open(opt.files1[lib].c_str(), nullptr), ok2 = b2.open(opt.files2[lib].c_str(), nullptr);
				bool same = ok1 && ok2 && a.fastq == b2.fastq; if (ok1) a.close(); if (ok2) b2.close();
				if (ok1 && ok2 && !same) { fprintf(stdout, "Error! %s and %s are with different format...\n", opt.files1[lib].c_str(), opt.files2[lib].c_str()); continue; }
				if (!ok1 || !ok2) continue;
			}
			if (!src.open(opt.files1[lib].c_str(), sep ? opt.files2[lib].c_str() : nullptr)) continue;
			while (true)
			{
				Job* j = free_q.take(); j->rb.clear();
				double ta = now_s();
				if (src.fill(j->rb, batch_reads, pair_end) <= 0) { free_q.put(j); break; }
				if (g_trace) fprintf(stderr, "[kart trace] read   %8d reads  %.3f..%.3f s\n", j->rb.n(), ta - g_t0, now_s() - g_t0);
				ready_q.put(j);
			}
			src.close();
		}
		pair_end_final = pair_end;
		ready_q.put(nullptr);
	});
	auto drain_reader = [&]() { while (Job* j = ready_q.take()) free_q.put(j); reader.join(); };

	kb_ctx_t* ctx = nullptr;
	int rc = kb_init(0, &ctx);
	if (g_trace) fprintf(stderr, "[kart trace] kb_init done %.3f s\n", now_s() - g_t0);
	if (rc) { fprintf(stderr, "Error! kart_b200 needs a CUDA device: %s\n", kb_strerror(rc)); drain_reader(); return 1; }
	host_cuda_ready();   // from here on the batch buffers are page-locked at allocation; the ones filled meanwhile are pinned in place below
	kb_index_host_t hi; idx.describe(&hi);
	// The full SA in HBM (8 bytes per BWT row, expanded on the device from the .sa samples) makes seeding several times cheaper per
	// read but costs seconds to expand on a multi-Gbp index (3.4 s at 3.1 Gbp) -- and this program is bound by parsing and SAM
	// text, not by the kernels. So by default the samples are expanded when that is cheap (up to 2^30 rows: under a second) or
	// when the input is long enough to pay for it; --full-sa / --sampled-sa decide explicitly.
	int expand = opt.expand_sa;
	if (expand == 2)
	{
		unsigned long long in_bytes = 0; struct stat fs;
		for (const auto& f : opt.files1) if (stat(f.c_str(), &fs) == 0) in_bytes += (unsigned long long)fs.st_size * (f.size() > 3 && f.compare(f.size() - 3, 3, ".gz") == 0 ? 4 : 1);
		for (const auto& f : opt.files2) if (stat(f.c_str(), &fs) == 0) in_bytes += (unsigned long long)fs.st_size * (f.size() > 3 && f.compare(f.size() - 3, 3, ".gz") == 0 ? 4 : 1);
		expand = (hi.seq_len <= (1ull << 30) || in_bytes >= (64ull << 30)) ? 2 : 0;
	}
	// page-locked result buffers of every rotating batch are allocated while the index goes to the device (r32 trace: 40-100 ms per
	// batch when the first pass through the rotation had to do it in the GPU worker)
	std::thread prealloc([&]() {
		for (size_t k = 0; k < jobs.size() && k < 4; k++)   // the first batches of the rotation; the others allocate on first use, under the running pipeline
		{
			Job& j = jobs[k];
			j.br.aln.reserve((size_t)batch_reads); j.br.pairs.reserve((size_t)batch_reads / 2 + 1); j.br.cigar.reserve((size_t)batch_reads * 4 + 1024);
		}
	});
	rc = kb_upload_index(ctx, &hi, expand);
	prealloc.join();
	if (rc != 0) { fprintf(stderr, "Error! index upload failed: %s (%s)\n", kb_strerror(rc), kb_last_error(ctx)); drain_reader(); kb_destroy(ctx); return 1; }

	if (g_trace) fprintf(stderr, "[kart trace] index uploaded %.3f s\n", now_s() - g_t0);
	SamSink out; BamWriter bam; const bool to_bam = opt.out_format == 1 && !opt.debug;
	std::vector<int32_t> name2id(idx.chr_name.size(), -1);   // bam_name2id: a repeated @SQ name keeps its first id (htslib sam.c:725-733)
	if (!opt.debug)
	{
		std::string h; sam_header(h, idx);
		if (to_bam)
		{
			if (!bam.open(opt.out_name.c_str(), opt.threads)) { fprintf(stderr, "Error! Cannot open file [%s]\n", opt.out_name.c_str()); drain_reader(); kb_destroy(ctx); exit(1); }
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
			if (!out.open(opt.out_name.c_str())) { fprintf(stderr, "Error! Cannot open file [%s]\n", opt.out_name.c_str()); drain_reader(); kb_destroy(ctx); exit(1); }
			out.write(h.data(), h.size());
		}
	}
	if (opt.silent) fprintf(stdout, "Start read mapping...\n");
	time_t t0 = time(NULL);
	long long total = 0, unmapped = 0, unique = 0, remapped = 0; PairState st;

	// stage 3: SAM/BAM text of batch k is formatted in slices while batch k+1 is on the GPU, and written (stage 4, its own thread) while
	// batch k+1 is being formatted: the batch's read and result buffers go back to the reader as soon as the text exists, and the
	// page-cache copy of ~390 MB per million reads (the long pole: 3-4 GB/s on ext4 whatever the number of writers, scripts/host_write_bench.cpp)
	// no longer waits for the formatter.
	struct OutBatch
	{
		std::vector<HBuf<char>> parts; std::vector<std::string> bparts; std::vector<std::vector<uint32_t>> rec_ends; int n = 0; double ta = 0, tb = 0;
		explicit OutBatch(int t) : parts(t), bparts(t), rec_ends(t) {}
	};
	OutBatch obuf0(io_threads), obuf1(io_threads); Channel<OutBatch*> ob_free, ob_ready;
	ob_free.put(&obuf0); ob_free.put(&obuf1);
	std::thread writer([&]() {
		while (Job* j = done_q.take())
		{
			OutBatch* ob = ob_free.take();
			std::vector<HBuf<char>>& parts = ob->parts; std::vector<std::string>& bparts = ob->bparts; std::vector<std::vector<uint32_t>>& rec_ends = ob->rec_ends;
			const ReadBatch& cur = j->rb; const int n = cur.n(), n_pe = j->n_pe; const bool fastq = cur.fastq;
			std::vector<long long> um(io_threads, 0), uq(io_threads, 0);
			ob->n = n; ob->ta = now_s();
			parallel_for(io_threads, (size_t)io_threads, [&](int, size_t t0_, size_t t1_) {
				for (size_t t = t0_; t < t1_; t++)
				{
					int lo = (int)((long long)n * (long long)t / io_threads), hi = (int)((long long)n * (long long)(t + 1) / io_threads);
					if (n_pe) { lo &= ~1; if ((int)t + 1 < io_threads) hi &= ~1; }
					parts[t].clear(); bparts[t].clear(); rec_ends[t].clear();
					// -m: cursors into the extra lines of the paired part and of the single-end tail
					auto first_of = [](const std::vector<kb_extra_t>& v, uint32_t r) { return (size_t)(std::lower_bound(v.begin(), v.end(), r, [](const kb_extra_t& e, uint32_t x) { return e.read < x; }) - v.begin()); };
					size_t xe = first_of(j->br.extra, (uint32_t)std::min(lo, n_pe)), xt = first_of(j->tail.extra, (uint32_t)std::max(lo - n_pe, 0));
					for (int r = lo; r < hi; r++)
					{
						bool in_pe = r < n_pe;
						const kb_aln_t& a = in_pe ? j->br.aln[r] : j->tail.aln[r - n_pe];
						const uint32_t* cg = in_pe ? j->br.cigar.data() : j->tail.cigar.data();
						if (a.score == 0) um[t]++; else if (a.mapq == 60) uq[t]++;
						if (to_bam) bam_read_record(bparts[t], rec_ends[t], name2id, cur, r, !(in_pe && (r & 1)), a, cg, fastq);
						else if (!opt.debug) sam_read_line(parts[t], idx, cur, r, !(in_pe && (r & 1)), a, cg, fastq);   // mate 2 of a mapped pair is held reverse-complemented
						const std::vector<kb_extra_t>& xv = in_pe ? j->br.extra : j->tail.extra; size_t& x = in_pe ? xe : xt; const uint32_t xr = (uint32_t)(in_pe ? r : r - n_pe);
						for (; x < xv.size() && xv[x].read == xr; x++)
						{
							if (to_bam) bam_read_record(bparts[t], rec_ends[t], name2id, cur, r, !(in_pe && (r & 1)), xv[x].aln, cg, fastq);
							else if (!opt.debug) sam_read_line(parts[t], idx, cur, r, !(in_pe && (r & 1)), xv[x].aln, cg, fastq);
						}
					}
				}
			});
			ob->tb = now_s();
			for (int t = 0; t < io_threads; t++) { unmapped += um[t]; unique += uq[t]; }
			free_q.put(j);
			ob_ready.put(ob);
		}
		ob_ready.put(nullptr);
	});
	std::thread sink([&]() {
		while (OutBatch* ob = ob_ready.take())
		{
			const double tc = now_s();
			if (to_bam) for (int t = 0; t < io_threads; t++) bam.append(ob->bparts[t], ob->rec_ends[t], true);
			else if (!opt.debug) out.write_parts(ob->parts);
			if (g_trace) fprintf(stderr, "[kart trace] format %8d reads  %.3f..%.3f s  write %.3f..%.3f s\n", ob->n, ob->ta - g_t0, ob->tb - g_t0, tc - g_t0, now_s() - g_t0);
			ob_free.put(ob);
		}
	});

	// stage 2: the GPUs
	g_multihit = opt.multihit; g_pack_threads = std::max(1, std::min(io_threads, 16));
	bool pair_end_seen = opt.pair_flag;
	std::mutex turn_m; std::condition_variable turn_cv; long long turn = 0;     // number of the batch that may settle next
	std::atomic<int> err{0}; std::atomic<int> est_pred{1500};
	Channel<Job*> work_q;
	auto gpu_worker = [&](int dev, kb_ctx_t* wctx) {
		PackBuf pk;
		kb_params_t pm; pm.min_seed_len = 0; pm.max_gaps = opt.max_gaps; pm.max_insert = 1500; pm.pacbio = opt.pacbio; pm.multihit = opt.multihit; pm.paired = 0;
		while (Job* j = work_q.take())
		{
			int wrc = err.load();
			ReadBatch& cur = j->rb; const bool pair_end = cur.pair_end;
			int n = cur.n(); double ta = now_s(), tb = ta;
			// a batch with an odd number of reads can only be the last one of its library: the reference sends a whole chunk through the
			// single-end branch when its read count is odd (Mapping.cpp:531,598), i.e. the final short chunk
			int n_pe = (!opt.pacbio && pair_end) ? ((n & 1) ? (n / chunk_reads) * chunk_reads : n) : 0;
			j->n_pe = n_pe;
			if (!wrc)
			{
				cur.seq.pin_now(); cur.seq_off.pin_now();
				if (g_trace && n > 100000) fprintf(stderr, "[kart trace]   pin %.1f ms\n", (now_s() - ta) * 1e3);
				if (n_pe > 0)
				{
					pm.paired = 1; kb_set_params(wctx, &pm);
					j->br.est_used.assign(n_pe / 2, est_pred.load());
					wrc = map_batch(wctx, cur.seq.data(), cur.seq_off.data(), n_pe, j->br.est_used.data(), j->br, pk);
				}
				if (!wrc && n > n_pe)
				{
					pm.paired = 0; kb_set_params(wctx, &pm);
					std::vector<uint64_t> off(cur.seq_off.data() + n_pe, cur.seq_off.data() + n + 1);
					uint64_t base = off[0]; for (auto& o : off) o -= base;
					wrc = map_batch(wctx, cur.seq.data() + base, off.data(), n - n_pe, nullptr, j->tail, pk);
				}
				tb = now_s();
			}
			// the recurrence, the counters and the hand-over to the writer happen in batch order
			{ std::unique_lock<std::mutex> g(turn_m); turn_cv.wait(g, [&]() { return turn == j->seq_no; }); }
			if (!wrc && !err.load() && n_pe > 0)
			{
				pm.paired = 1; kb_set_params(wctx, &pm);
				wrc = settle_est(wctx, cur, j->br, st, chunk_reads, n_pe, &remapped, pk);
				est_pred.store(est_of(st));
			}
			if (wrc && !err.load())
			{
				err.store(wrc);
				fprintf(stderr, "\nError! GPU mapping failed on device %d: %s (%s)\n", dev, kb_strerror(wrc), kb_last_error(wctx));
			}
			if (!err.load())
			{
				pair_end_seen = pair_end;
				if (!opt.silent) { fprintf(stdout, "\r%lld %s reads have been processed in %ld seconds...", total, pair_end ? "paired-end" : "singled-end", (long)(time(NULL) - t0)); fflush(stdout); }
				total += n;
				if (g_trace) fprintf(stderr, "[kart trace] gpu%d   %8d reads  %.3f..%.3f s  settle ..%.3f s (remapped %lld)\n", dev, n, ta - g_t0, tb - g_t0, now_s() - g_t0, remapped);
				done_q.put(j);
			}
			else free_q.put(j);
			{ std::lock_guard<std::mutex> g(turn_m); turn++; }
			turn_cv.notify_all();
		}
		work_q.put(nullptr);   // pass the end marker on to the next worker
	};
	std::vector<std::thread> extra_workers;
	std::thread worker0(gpu_worker, 0, ctx);
	{
		// dispatcher: numbers the batches; a further device is brought up for every two batches beyond what the running ones took
		long long seq_no = 0; int n_dev = 1, n_visible = 1;
		if (n_dev_max > 1) n_visible = std::min(n_dev_max, kb_device_count());
		while (Job* j = ready_q.take())
		{
			j->seq_no = seq_no++;
			work_q.put(j);
			if (n_dev < n_visible && seq_no >= 2 * (long long)n_dev && !err.load())
			{
				const int dev = n_dev++;
				extra_workers.emplace_back([&, dev]() {
					kb_ctx_t* c2 = nullptr; double ta = now_s();
					int r2 = kb_init(dev, &c2);
					if (!r2) r2 = kb_clone_index(c2, ctx);
					if (r2) { fprintf(stderr, "Warning! device %d is not used: %s (%s)\n", dev, kb_strerror(r2), c2 ? kb_last_error(c2) : ""); if (c2) kb_destroy(c2); return; }
					if (g_trace) fprintf(stderr, "[kart trace] device %d up (context + index clone) %.3f..%.3f s\n", dev, ta - g_t0, now_s() - g_t0);
					gpu_worker(dev, c2);
				});
			}
		}
		work_q.put(nullptr);
	}
	worker0.join();
	for (auto& t : extra_workers) t.join();
	rc = err.load();
	reader.join();
	done_q.put(nullptr); writer.join(); sink.join();
	const bool pair_end = opt.files1.empty() ? pair_end_seen : pair_end_final;
	fprintf(stdout, "\rAll the %lld %s reads have been processed in %lld seconds.\n", total, pair_end ? "paired-end" : "single-end", (long long)(time(NULL) - t0));
	out.close();
	if (to_bam) bam.close();
	if (total > 0)
	{
		if (pair_end) fprintf(stdout, "\t# of total mapped sequences = %lld (sensitivity = %.2f%%)\n\t# of paired sequences = %lld (%.2f%%), average insert size = %d\n", total - unmapped, (int)(10000 * (1.0 * (total - unmapped) / total) + 0.5) / 100.0, st.iPaired, (int)(10000 * (1.0 * st.iPaired / total) + 0.5) / 100.0, (st.iPaired > 1 ? (int)(st.iDistance / (st.iPaired >> 1)) : 0));
		else fprintf(stdout, "\t# of total mapped sequences = %lld (sensitivity = %.2f%%)\n", total - unmapped, (int)(10000 * (1.0 * (total - unmapped) / total) + 0.5) / 100.0);
		fprintf(stdout, "Alignment output: %s\n", opt.out_name.c_str());
	}
	(void)unique; (void)remapped;
	// Everything is on disk. Tearing the CUDA context and gigabytes of page-locked buffers down costs about half a second that
	// buys nothing at process exit (KART_B200_CLEAN_EXIT=1 keeps the orderly path for leak checkers).
	if (!getenv("KART_B200_CLEAN_EXIT")) { fflush(stdout); fflush(stderr); _exit(rc ? 1 : 0); }
	kb_destroy(ctx);
	return rc ? 1 : 0;
}
