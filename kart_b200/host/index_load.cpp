// BWA-format index files -> host arrays (formats: reference src/bwt_index.cpp:16-36 .sa, :103-122 .bwt, :38-90 .ann/.amb, :230-259 .pac)
#include "kart_host.h"
#include <fstream>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

// The big files (.bwt, .sa, .pac: 5.4 GB for a 3.1 Gbp genome) are read once, by several threads, into huge-page backed memory and
// handed to kb_upload_index as they are (the sa[0] = -1 entry of bwt_restore_sa is put in front of the samples in place). Start-up
// is what a short run consists of: zero-filled vectors + fread + a second copy cost seconds here (r18: 8.5 s before the first read was
// mapped), a populated private mapping still 2.3 s of page faults (r19); parallel pread runs at memory speed.
#include <thread>
static bool read_file(const std::string& fn, HostIndex::Mapping& m)
{
	int fd = open(fn.c_str(), O_RDONLY); if (fd < 0) return false;
	struct stat st; if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return false; }
	const size_t n = (size_t)st.st_size, huge = (size_t)2 << 20, cap = (n + huge - 1) / huge * huge;
	void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (p == MAP_FAILED) { close(fd); return false; }
	madvise(p, cap, MADV_HUGEPAGE);
	unsigned hw = std::thread::hardware_concurrency(); int nt = (int)(hw ? (hw > 16 ? 16 : hw) : 4);
	if (n < ((size_t)64 << 20)) nt = 1;
	std::vector<std::thread> th; std::vector<char> ok((size_t)nt, 1);
	auto work = [&](int t) {
		size_t lo = n / nt * t, hi = t + 1 == nt ? n : n / nt * (t + 1);
		while (lo < hi) { ssize_t r = pread(fd, (uint8_t*)p + lo, hi - lo > ((size_t)1 << 30) ? ((size_t)1 << 30) : hi - lo, (off_t)lo); if (r <= 0) { ok[(size_t)t] = 0; return; } lo += (size_t)r; }
	};
	for (int t = 1; t < nt; t++) th.emplace_back(work, t);
	work(0); for (auto& x : th) x.join();
	close(fd);
	for (char c : ok) if (!c) { munmap(p, cap); return false; }
	m.p = (uint8_t*)p; m.n = n; m.cap = cap;
	return true;
}
HostIndex::~HostIndex() { for (Mapping* m : {&map_bwt, &map_sa, &map_pac}) if (m->p) munmap(m->p, m->cap); }

bool check_index_files(const std::string& prefix)
{
	const char* ext[3] = {".ann", ".amb", ".pac"};
	for (int i = 0; i < 3; i++) { std::ifstream f((prefix + ext[i]).c_str()); if (!f.is_open()) return false; }
	return true;
}

bool HostIndex::load(const std::string& prefix, std::string& err)
{
	if (!read_file(prefix + ".bwt", map_bwt) || map_bwt.n < 40) { err = "cannot read " + prefix + ".bwt"; return false; }
	memcpy(&primary, map_bwt.p, 8); memcpy(&L2[1], map_bwt.p + 8, 32); L2[0] = 0; seq_len = L2[4];
	bwt = (const uint32_t*)(map_bwt.p + 40); bwt_words = (map_bwt.n - 40) / 4;
	if (!read_file(prefix + ".sa", map_sa) || map_sa.n < 56) { err = "cannot read " + prefix + ".sa"; return false; }
	uint64_t intv; memcpy(&intv, map_sa.p + 40, 8); sa_intv = (int)intv;
	if (sa_intv <= 0) { err = "bad SA interval"; return false; }
	n_sa = (seq_len + sa_intv) / sa_intv;
	if (map_sa.n - 56 < (n_sa - 1) * 8) { err = prefix + ".sa is too short"; return false; }
	// header word 6 (the text length) sits right in front of the samples: it becomes sa[0] = -1 in this private mapping
	uint64_t minus1 = (uint64_t)-1; memcpy(map_sa.p + 48, &minus1, 8);
	sa = (const uint64_t*)(map_sa.p + 48);
	FILE* fp = fopen((prefix + ".ann").c_str(), "r"); if (!fp) { err = "cannot read " + prefix + ".ann"; return false; }
	long long lp; int nseq; unsigned seed;
	if (fscanf(fp, "%lld%d%u", &lp, &nseq, &seed) != 3 || nseq <= 0) { fclose(fp); err = "bad .ann"; return false; }
	l_pac = lp;
	for (int i = 0; i < nseq; i++)
	{
		unsigned gi; char name[1024]; long long off; int len, namb, c;
		if (fscanf(fp, "%u%1023s", &gi, name) != 2) { fclose(fp); err = "bad .ann"; return false; }
		while ((c = fgetc(fp)) != '\n' && c != EOF) {}
		if (fscanf(fp, "%lld%d%d", &off, &len, &namb) != 3) { fclose(fp); err = "bad .ann"; return false; }
		chr_name.push_back(name); chr_len.push_back(len);
	}
	fclose(fp);
	if (!read_file(prefix + ".pac", map_pac)) { err = "cannot read " + prefix + ".pac"; return false; }
	const size_t need = (size_t)(l_pac / 4 + 1);
	if (map_pac.n >= need) pac = map_pac.p;
	else { pac_copy.assign(need, 0); memcpy(pac_copy.data(), map_pac.p, map_pac.n); pac = pac_copy.data(); }
	return true;
}

void HostIndex::describe(kb_index_host_t* o) const
{
	o->primary = primary; for (int i = 0; i < 5; i++) o->L2[i] = L2[i]; o->seq_len = seq_len;
	o->bwt = bwt; o->bwt_words = bwt_words; o->sa = sa; o->n_sa = n_sa; o->sa_intv = sa_intv;
	o->pac = pac; o->l_pac = l_pac; o->n_chr = (int)chr_len.size(); o->chr_len = chr_len.data();
}
