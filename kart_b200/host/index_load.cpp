// BWA-format index files -> host arrays (formats: reference src/bwt_index.cpp:16-36 .sa, :103-122 .bwt, :38-90 .ann/.amb, :230-259 .pac)
#include "kart_host.h"
#include <fstream>
#include <string.h>

static bool slurp(const std::string& fn, std::vector<uint8_t>& buf)
{
	FILE* fp = fopen(fn.c_str(), "rb"); if (!fp) return false;
	fseek(fp, 0, SEEK_END); long n = ftell(fp); fseek(fp, 0, SEEK_SET);
	buf.resize(n > 0 ? n : 0);
	size_t got = n > 0 ? fread(buf.data(), 1, n, fp) : 0; fclose(fp);
	return (long)got == n;
}

bool check_index_files(const std::string& prefix)
{
	const char* ext[3] = {".ann", ".amb", ".pac"};
	for (int i = 0; i < 3; i++) { std::ifstream f((prefix + ext[i]).c_str()); if (!f.is_open()) return false; }
	return true;
}

bool HostIndex::load(const std::string& prefix, std::string& err)
{
	std::vector<uint8_t> raw;
	if (!slurp(prefix + ".bwt", raw) || raw.size() < 40) { err = "cannot read " + prefix + ".bwt"; return false; }
	memcpy(&primary, raw.data(), 8); memcpy(&L2[1], raw.data() + 8, 32); L2[0] = 0; seq_len = L2[4];
	bwt.resize((raw.size() - 40) / 4); memcpy(bwt.data(), raw.data() + 40, bwt.size() * 4);
	if (!slurp(prefix + ".sa", raw) || raw.size() < 56) { err = "cannot read " + prefix + ".sa"; return false; }
	uint64_t intv; memcpy(&intv, raw.data() + 40, 8); sa_intv = (int)intv;
	if (sa_intv <= 0) { err = "bad SA interval"; return false; }
	uint64_t n_sa = (seq_len + sa_intv) / sa_intv;
	sa.assign(n_sa, 0); sa[0] = (uint64_t)-1;
	size_t body = raw.size() - 56; if (body > (n_sa - 1) * 8) body = (n_sa - 1) * 8;
	memcpy(sa.data() + 1, raw.data() + 56, body);
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
	if (!slurp(prefix + ".pac", raw)) { err = "cannot read " + prefix + ".pac"; return false; }
	pac.assign((size_t)(l_pac / 4 + 1), 0);
	memcpy(pac.data(), raw.data(), raw.size() < pac.size() ? raw.size() : pac.size());
	return true;
}

void HostIndex::describe(kb_index_host_t* o) const
{
	o->primary = primary; for (int i = 0; i < 5; i++) o->L2[i] = L2[i]; o->seq_len = seq_len;
	o->bwt = bwt.data(); o->bwt_words = bwt.size(); o->sa = sa.data(); o->n_sa = sa.size(); o->sa_intv = sa_intv;
	o->pac = pac.data(); o->l_pac = l_pac; o->n_chr = (int)chr_len.size(); o->chr_len = chr_len.data();
}
