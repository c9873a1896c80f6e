// BWA-format index files -> host arrays (formats: reference src/bwt_index.cpp:16-36 .sa, :103-122 .bwt, :38-90 .ann/.amb, :230-259 .pac)
#include "kart_host.h"
#include <fstream>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

// The big files (.bwt, .sa, .pac: 5.4 GB for a 3.1 Gbp genome) are mapped, not copied: the only consumer is kb_upload_index, which
// sends them to the device once. A private mapping lets the loader put the sa[0] = -1 entry (bwt_restore_sa) in front of the
// samples in place. Start-up is what a short run consists of: reading into zero-filled vectors and copying again cost more than
// mapping the reads of a 1 M-read job.
static bool map_file(const std::string& fn, HostIndex::Mapping& m)
{
	int fd = open(fn.c_str(), O_RDONLY); if (fd < 0) return false;
	struct stat st; if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return false; }
	void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_POPULATE, fd, 0);
	close(fd);
	if (p == MAP_FAILED) return false;
	m.p = (uint8_t*)p; m.n = (size_t)st.st_size;
	return true;
}
HostIndex::~HostIndex() { for (Mapping* m : {&map_bwt, &map_sa, &map_pac}) if (m->p) munmap(m->p, m->n); }

bool check_index_files(const std::string& prefix)
{
	const char* ext[3] = {".ann", ".amb", ".pac"};
	for (int i = 0; i < 3; i++) { std::ifstream f((prefix + ext[i]).c_str()); if (!f.is_open()) return false; }
	return true;
}

bool HostIndex::load(const std::string& prefix, std::string& err)
{
	if (!map_file(prefix + ".bwt", map_bwt) || map_bwt.n < 40) { err = "cannot read " + prefix + ".bwt"; return false; }
	memcpy(&primary, map_bwt.p, 8); memcpy(&L2[1], map_bwt.p + 8, 32); L2[0] = 0; seq_len = L2[4];
	bwt = (const uint32_t*)(map_bwt.p + 40); bwt_words = (map_bwt.n - 40) / 4;
	if (!map_file(prefix + ".sa", map_sa) || map_sa.n < 56) { err = "cannot read " + prefix + ".sa"; return false; }
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
	if (!map_file(prefix + ".pac", map_pac)) { err = "cannot read " + prefix + ".pac"; return false; }
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
