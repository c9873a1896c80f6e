// Page-cache write throughput for the CLI's output stage: 390 MB batches (one million 150-bp SAM records) written by (0) T parallel pwrite calls,
// (1) T threads copying into an mmap of the file, (2) one write. Usage: host_write_bench <file> <mode 0|1|2> <threads> <batches>
// profiles/r32_host_write.txt: ~3-3.5 GB/s on the build container's ext4 whatever the mode -- the long pole of kart_b200/bin/kart (SAM out).
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <time.h>
static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv)
{
	const char* path = argv[1]; int mode = atoi(argv[2]), T = atoi(argv[3]); int batches = atoi(argv[4]);
	size_t B = 390u << 20; char* src = (char*)malloc(B); memset(src, 'A', B);
	int fd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0666); off_t at = 0;
	double t0 = now();
	for (int b = 0; b < batches; b++)
	{
		size_t part = B / T;
		if (mode == 0) { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back([&, t]() { size_t lo = t * part, hi = t + 1 == T ? B : lo + part; off_t o = at + lo; const char* p = src + lo; size_t n = hi - lo; while (n) { ssize_t w = pwrite(fd, p, n > (8u << 20) ? (8u << 20) : n, o); if (w <= 0) exit(1); p += w; n -= w; o += w; } }); for (auto& t : th) t.join(); }
		else if (mode == 1)
		{
			if (ftruncate(fd, at + B)) exit(2);
			off_t a0 = at & ~(off_t)4095; size_t len = (size_t)(at + B - a0);
			char* m = (char*)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, a0); if (m == MAP_FAILED) exit(3);
			char* dst = m + (at - a0);
			std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back([&, t]() { size_t lo = t * part, hi = t + 1 == T ? B : lo + part; memcpy(dst + lo, src + lo, hi - lo); }); for (auto& t : th) t.join();
			munmap(m, len);
		}
		else { const char* p = src; size_t n = B; while (n) { ssize_t w = write(fd, p, n); if (w <= 0) exit(1); p += w; n -= w; } }
		at += B;
	}
	double t1 = now();
	close(fd);
	printf("mode %d threads %d: %.2f GB/s (%.0f ms per 390 MB batch), close %.3f s\n", mode, T, batches * (double)B / (t1 - t0) / 1e9, (t1 - t0) / batches * 1e3, now() - t1);
	unlink(path);
}
