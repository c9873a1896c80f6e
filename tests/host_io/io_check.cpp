// TEST INFRASTRUCTURE. Dumps the batches ReadSource produces for a set of input files, through the block path (`blocks`) or
// the entry-at-a-time path (`serial`), so that tests/test_host_io.py can require both to yield the same reads, names,
// qualities and batch boundaries.  Usage: io_check blocks|serial <threads> <batch_reads> <pair_end 0|1> file1 [file2]
#include "../../kart_b200/host/kart_host.h"
#include <stdlib.h>
extern "C" void* kb_host_alloc(uint64_t n) { return malloc(n ? n : 1); }
extern "C" void kb_host_free(void* p) { free(p); }
extern "C" int kb_host_register(void*, uint64_t) { return 0; }
extern "C" void kb_host_unregister(void*) {}
int main(int argc, char** argv)
{
	if (argc < 6) return 2;
	bool blocks = std::string(argv[1]) == "blocks"; int threads = atoi(argv[2]), batch = atoi(argv[3]); bool pe = atoi(argv[4]) != 0;
	ReadSource src; src.threads = threads;
	if (!src.open(argv[5], argc > 6 ? argv[6] : nullptr)) { printf("open failed\n"); return 1; }
	ReadBatch b; int guard = 0;
	while (guard++ < 100000)
	{
		b.clear();
		int got = blocks ? src.fill(b, batch, pe) : src.fill_serial(b, batch, pe);
		if (got <= 0) break;
		printf("batch %d fastq %d\n", got, (int)src.fastq);
		for (int r = 0; r < b.n(); r++)
		{
			size_t L = b.seq_off[r + 1] - b.seq_off[r];
			printf("[%.*s]\t", (int)(b.name_off[r + 1] - b.name_off[r]), b.names.data() + b.name_off[r]);
			fwrite(b.seq.data() + b.seq_off[r], 1, L, stdout); fputc('\t', stdout);
			if (src.fastq) fwrite(b.qual.data() + b.seq_off[r], 1, L, stdout);
			fputc('\n', stdout);
		}
	}
	src.close();
	return 0;
}
