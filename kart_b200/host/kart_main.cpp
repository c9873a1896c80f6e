// Command line of kart_b200: same flags, defaults, messages and exit codes as the reference's src/main.cpp:87-214.
// `kart index` builds the BWA-format files itself (index_build.cpp); `kart update` is not supported.
#include "kart_host.h"
#include <thread>
#include <algorithm>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

static const char* kVersion = "2.5.6";

static void usage(const char* program)
{
	fprintf(stdout, "kart v%s (Hsin-Nan Lin & Wen-Lian Hsu)\n\n", kVersion);
	fprintf(stdout, "Usage: %s -i Index_Prefix -f <ReadFile_A1 ReadFile_B1 ...> [-f2 <ReadFile_A2 ReadFile_B2 ...>] -o Output\n\n", program);
	fprintf(stdout, "Options: -t INT        number of threads [4]\n");
	fprintf(stdout, "         -f            files with #1 mates reads (format:fa, fq, fq.gz)\n");
	fprintf(stdout, "         -f2           files with #2 mates reads (format:fa, fq, fq.gz)\n");
	fprintf(stdout, "         -o            alignment filename in SAM format [output.sam]\n");
	fprintf(stdout, "         -bo           alignment filename in BAM format\n");
	fprintf(stdout, "         -m            output multiple alignments\n");
	fprintf(stdout, "         -g INT        max gaps (indels) [5]\n");
	fprintf(stdout, "         -p            paired-end reads are interlaced in the same file\n");
	fprintf(stdout, "         -pacbio       pacbio data\n");
	fprintf(stdout, "         -v            version\n");
	fprintf(stdout, "\n");
}

static bool check_output_name(const std::string& name)   // CheckOutputFileName, main.cpp:33
{
	bool ok = true;
	if (name != "output.sam")
	{
		struct stat s;
		if (stat(name.c_str(), &s) == 0 && (s.st_mode & S_IFDIR)) { ok = false; fprintf(stdout, "Warning: %s is a directory!\n", name.c_str()); }
		for (size_t i = 0; i < name.size(); i++)
		{
			char c = name[i];
			if (!(isalnum((unsigned char)c) || c == '/' || c == '.' || c == '-' || c == '_')) { ok = false; fprintf(stdout, "Warning: [%s] is not a valid filename!\n", name.c_str()); break; }
		}
	}
	return ok;
}

static bool check_inputs(const RunOptions& o)   // CheckInputFiles, main.cpp:63
{
	struct stat s; bool ok = true;
	for (const auto& f : o.files1) if (stat(f.c_str(), &s) == -1) { ok = false; fprintf(stdout, "Cannot access file:[%s]\n", f.c_str()); }
	for (const auto& f : o.files2) if (stat(f.c_str(), &s) == -1) { ok = false; fprintf(stdout, "Cannot access file:[%s]\n", f.c_str()); }
	return ok;
}

int main(int argc, char* argv[])
{
	RunOptions o;
	if (argc == 1 || strcmp(argv[1], "-h") == 0) { usage(argv[0]); return 0; }
	if (strcmp(argv[1], "update") == 0) { fprintf(stderr, "kart_b200: `update` is not supported\n"); return 0; }
	if (strcmp(argv[1], "index") == 0)   // main.cpp:112-120
	{
		if (argc == 4) return build_index(argv[2], argv[3], (int)std::max(1u, std::thread::hardware_concurrency()));
		if (argc == 5 && strcmp(argv[2], "-gpu") == 0 && strcmp(argv[3], "-pac") == 0) return build_index_gpu(nullptr, argv[4], true);
		if (argc == 5 && strcmp(argv[2], "-gpu") == 0) return build_index_gpu(argv[3], argv[4], false);
		fprintf(stderr, "usage: %s index ref.fa prefix\n       %s index -gpu ref.fa prefix    (BWT, Occ and SA built on the GPU; same files)\n       %s index -gpu -pac prefix      (the same from an existing prefix.pac + prefix.ann)\n", argv[0], argv[0], argv[0]);
		return 0;
	}
	for (int i = 1; i < argc; i++)
	{
		std::string p = argv[i];
		if (p == "-i") o.index_prefix = (++i < argc) ? argv[i] : "";
		else if (p == "-f") { while (++i < argc && argv[i][0] != '-') o.files1.push_back(argv[i]); i--; }
		else if (p == "-f2") { while (++i < argc && argv[i][0] != '-') o.files2.push_back(argv[i]); i--; }
		else if (p == "-t" && i + 1 < argc) { if ((o.threads = atoi(argv[++i])) <= 0) { fprintf(stdout, "Warning! Thread number should be a positive number!\n"); o.threads = 4; } }
		else if (p == "-g") { if (++i < argc && (o.max_gaps = atoi(argv[i])) < 0) o.max_gaps = 0; }
		else if (p == "-o") { o.out_format = 0; if (++i < argc) o.out_name = argv[i]; }
		else if (p == "-bo") { o.out_format = 1; if (++i < argc) o.out_name = argv[i]; }
		else if (p == "-silent") o.silent = true;
		else if (p == "-pacbio") o.pacbio = true;
		else if (p == "-m") o.multihit = true;
		else if (p == "-pair" || p == "-p") o.pair_flag = true;
		else if (p == "-d" || p == "-debug") o.debug = true;
		else if (p == "-v" || p == "--version") { fprintf(stdout, "kart v%s\n\n", kVersion); exit(0); }
		else if (p == "--batch" && i + 1 < argc) o.batch_reads = atoi(argv[++i]);        // kart_b200 extension: reads per GPU batch
		else if (p == "--gpus" && i + 1 < argc) o.n_gpus = atoi(argv[++i]);              // kart_b200 extension: devices to use (default: all visible)
		else if (p == "--full-sa") o.expand_sa = 1;                                      // kart_b200 extension: expand the SA in HBM (default: when it fits)
		else if (p == "--sampled-sa") o.expand_sa = 0;                                   // kart_b200 extension: keep the .sa sampling on the device
		else { fprintf(stdout, "Error! Unknown parameter: %s\n", argv[i]); usage(argv[0]); exit(1); }
	}
	if (o.files1.empty()) { fprintf(stdout, "Error! Please specify a valid read input!\n"); usage(argv[0]); exit(1); }
	if (!o.files2.empty() && o.files1.size() != o.files2.size())
	{
		fprintf(stdout, "Error! Paired-end reads input numbers do not match!\n");
		fprintf(stdout, "Read1:\n"); for (const auto& f : o.files1) fprintf(stdout, "\t%s\n", f.c_str());
		fprintf(stdout, "Read2:\n"); for (const auto& f : o.files2) fprintf(stdout, "\t%s\n", f.c_str());
		exit(1);
	}
	if (!check_inputs(o) || !check_output_name(o.out_name)) exit(0);
	HostIndex idx; std::string err;
	if (o.index_prefix.empty() || !check_index_files(o.index_prefix)) { fprintf(stdout, "Error! Please specify a valid reference index!\n"); usage(argv[0]); exit(1); }
	// creating the CUDA context takes about a second on a B200 box: do it while the index files are read (kb_host_alloc
	// initialises the runtime from any thread, before kb_init)
	std::thread cuda_warm([]() { kb_host_free(kb_host_alloc(1)); });
	fprintf(stdout, "Load the genome index files...");
	bool ok = idx.load(o.index_prefix, err);
	fprintf(stdout, "\n");
	if (!ok) { cuda_warm.join(); fprintf(stdout, "\n\nError! Index files are corrupt!\n"); exit(1); }
	cuda_warm.detach();
	fprintf(stdout, "Load the reference sequences...\n");
	return run_mapping(o, idx);
}
