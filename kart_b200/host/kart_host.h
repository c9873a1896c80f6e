// Host side of kart_b200: freshly written C++ with the observable behaviour of the reference's host code
// (CLI src/main.cpp, index loading src/bwt_index.cpp, read input src/GetData.cpp, chunk loop + SAM text src/Mapping.cpp).
// All per-read computation goes through the C ABI in include/kart_b200.h; nothing here maps reads on the CPU.
#ifndef KART_HOST_H
#define KART_HOST_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <functional>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/kart_b200.h"

struct HostIndex
{
	struct Mapping { uint8_t* p = nullptr; size_t n = 0, cap = 0; };
	Mapping map_bwt, map_sa, map_pac;                              // the files' bytes in huge-page backed memory (index_load.cpp)
	const uint32_t* bwt = nullptr; uint64_t bwt_words = 0; const uint64_t* sa = nullptr; uint64_t n_sa = 0; const uint8_t* pac = nullptr;
	std::vector<uint8_t> pac_copy;                                 // only when the .pac file is shorter than l_pac / 4 + 1 bytes
	HostIndex() = default; HostIndex(const HostIndex&) = delete; HostIndex& operator=(const HostIndex&) = delete; ~HostIndex();
	std::vector<std::string> chr_name; std::vector<int64_t> chr_len;
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0; int sa_intv = 32; int64_t l_pac = 0;
	bool load(const std::string& prefix, std::string& err);       // bwa_idx_load + RestoreReferenceInfo
	void describe(kb_index_host_t* out) const;
};
int build_index(const char* fasta, const char* prefix, int threads);
int build_index_gpu(const char* fasta, const char* prefix, bool from_pac);   // `kart index -gpu [-pac]`: BWT / Occ / SA built on the device (csrc/kb_index_build.cu), same bytes   // `kart index`: BWA-format files, byte-identical to the reference builder's (index_build.cpp)
bool check_index_files(const std::string& prefix);                // CheckBWAIndexFiles, GetData.cpp:222

// Grow-only host array without value-initialisation. `pinned` arrays come from kb_host_alloc (page-locked, so the copies of
// kb_map_chunk are asynchronous DMA at full PCIe rate) and fall back to malloc when that fails. Until the CUDA context
// exists (host_cuda_ready) they are plain page-aligned memory, so that the reader can fill the first batches while the device
// is being initialised; pin_now() page-locks such an array in place once the context is up.
void* host_buf_alloc(size_t bytes, bool want_pinned, bool* got_pinned);
void  host_buf_free(void* p, bool pinned);
void  host_cuda_ready();
template <class T> struct HBuf
{
	T* p = nullptr; size_t n = 0, cap = 0; bool want_pinned = false, is_pinned = false, registered = false;
	explicit HBuf(bool pin = false) : want_pinned(pin) {}
	HBuf(const HBuf&) = delete; HBuf& operator=(const HBuf&) = delete;
	~HBuf() { release(); }
	void release() { if (p) { if (registered) kb_host_unregister(p); host_buf_free(p, is_pinned); } p = nullptr; registered = false; }
	void pin_now() { if (want_pinned && !is_pinned && !registered && p && cap) registered = kb_host_register(p, cap * sizeof(T)) == KB_OK; }
	void reserve(size_t c)
	{
		if (c <= cap) return;
		size_t nc = cap + cap / 2; if (nc < c) nc = c; if (nc < 64) nc = 64;
		bool pin = false; T* q = (T*)host_buf_alloc(nc * sizeof(T), want_pinned, &pin);
		if (n) memcpy(q, p, n * sizeof(T));
		release();
		p = q; cap = nc; is_pinned = pin;
	}
	void resize(size_t k) { reserve(k); n = k; }                       // new elements are uninitialised
	void clear() { n = 0; }
	void push_back(const T& v) { if (n == cap) reserve(n + 1); p[n++] = v; }
	void append(const T* a, size_t k) { reserve(n + k); memcpy(p + n, a, k * sizeof(T)); n += k; }
	void assign(size_t k, const T& v) { resize(k); for (size_t i = 0; i < k; i++) p[i] = v; }
	T* data() { return p; } const T* data() const { return p; }
	size_t size() const { return n; } bool empty() const { return n == 0; }
	T& operator[](size_t i) { return p[i]; } const T& operator[](size_t i) const { return p[i]; }
	T& back() { return p[n - 1]; }
};

// One batch of reads in structure-of-arrays form (what the C ABI takes) plus what SAM output needs.
struct ReadBatch
{
	HBuf<uint8_t> seq{true};  HBuf<uint64_t> seq_off{true};           // mate 2 stored reverse-complemented (GetData.cpp:125-135)
	HBuf<char> qual;                                                   // same offsets as seq (mate 2 reversed), empty for FASTA
	HBuf<char> names;   HBuf<uint32_t> name_off;
	bool fastq = true, pair_end = false;                               // format / pairing of the library this batch was read from
	int n() const { return (int)seq_off.size() - 1; }
	void clear() { seq.clear(); seq_off.assign(1, 0); qual.clear(); names.clear(); name_off.assign(1, 0); }
};

// One input file behind a gzread-like call. Plain and ordinary gzip files go through zlib's gzFile (one inflate stream: a gzip
// stream has no block boundaries to split at). A file that is a chain of BGZF blocks (bgzip / htslib: self-contained deflate members
// of at most 64 KB, each announcing its compressed size in a "BC" extra field) is mapped, and read() inflates as many whole blocks as
// fit the request on `threads` workers, checking every block's CRC-32 and length like gzread does (the reference reads .gz input through
// gzgets on one thread, src/GetData.cpp:184-219).
struct GzInput
{
	gzFile fp = nullptr;
	const uint8_t* map = nullptr; size_t map_len = 0, at = 0; std::vector<char> carry; size_t carry_pos = 0; int threads = 1; bool bgzf = false, failed = false;
	bool open(const char* path, int n_threads);
	int read(char* dst, unsigned n);      // bytes delivered (possibly fewer than n), 0 at the end of the input, < 0 on a damaged file
	void close();
};

// FASTA / FASTQ (plain or .gz) input with the reference's parsing rules. FASTQ goes through a block reader: every stream
// buffers a large piece of (decompressed) text, newline positions are indexed by `threads` workers, and the records are
// parsed and copied into the batch by the same workers (read_input.cpp). FASTA keeps the entry-at-a-time path.
class ReadSource   // GetNextChunk / gzGetNextChunk semantics on top of zlib (which reads plain files transparently)
{
public:
	bool open(const char* f1, const char* f2);
	void close();
	bool fastq = true; int threads = 1;
	// appends up to max_reads reads (always whole pairs of entries, like the reference) ; returns reads appended
	int fill(ReadBatch& b, int max_reads, bool pair_end);
	int fill_serial(ReadBatch& b, int max_reads, bool pair_end);       // entry-at-a-time path (FASTA; and the cross-check of the block path in tests)
	struct Stream
	{
		GzInput in; std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false; std::string pending; bool has_pending = false;
		// block path: text[lo, hi) is buffered, nl[] holds the offsets of the newlines in it, rec = lines already consumed
		HBuf<char> text; size_t lo = 0, hi = 0; HBuf<uint32_t> nl; size_t nl_used = 0; bool drained = false;
	};
private:
	Stream s1, s2; bool two = false;
	bool line(Stream& s, std::string& out);
	bool entry(Stream& s, std::string& name, std::string& seq, std::string& qual);
	size_t buffer_records(Stream& s, size_t want);                     // ensures `want` complete 4-line records (or everything up to EOF) are indexed; returns the count
	int fill_blocks(ReadBatch& b, int max_reads, bool pair_end);
};

void parallel_for(int threads, size_t n, const std::function<void(int, size_t, size_t)>& body);   // body(worker, lo, hi) over [0,n) in `threads` contiguous slices

struct RunOptions
{
	std::string index_prefix, out_name = "output.sam";
	std::vector<std::string> files1, files2;
	int threads = 4, max_gaps = 5, out_format = 0, n_gpus = 0 /* 0: all visible devices (--gpus N) */; bool pair_flag = false, pacbio = false, multihit = false, silent = false, debug = false;
	int batch_reads = 1 << 20 /* large enough for kb_map_chunk's slot pipeline (>= 262 144 reads) to overlap copies and kernels */; int expand_sa = 2;   // 0 sampled SA, 1 full SA in HBM, 2 full SA when the device has room (kb_upload_index)
};

int run_mapping(const RunOptions& opt, const HostIndex& idx);     // Mapping(), src/Mapping.cpp:639

// SAM text (src/Mapping.cpp:177-315)
void sam_header(std::string& out, const HostIndex& idx);
void sam_read_line(HBuf<char>& out, const HostIndex& idx, const ReadBatch& b, int r, bool stored_fwd, const kb_aln_t& a, const uint32_t* cigar, bool fastq);


// BAM output (src/Mapping.cpp:610-621 via htslib sam_parse1 + sam_write1), see bam_writer.cpp
void bam_tables_init();
void bam_read_record(std::string& out, std::vector<uint32_t>& rec_end, const std::vector<int32_t>& name2id, const ReadBatch& b, int r, bool stored_fwd,
                     const kb_aln_t& a, const uint32_t* cigar, bool fastq);
class BamWriter
{
public:
	bool open(const char* path, int n_threads);
	void write_header(const std::string& text, const std::vector<std::string>& names, const std::vector<int64_t>& lens);
	void append(const std::string& bytes, const std::vector<uint32_t>& unit_end, bool records);   // unit_end: end offset of every record in `bytes`
	bool close();
private:
	void emit(const std::string& stream, const std::vector<size_t>& cuts);
	void flush();
	FILE* fp = nullptr; int threads = 1; std::string pend;
};

#endif
