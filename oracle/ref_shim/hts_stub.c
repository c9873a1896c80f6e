/* TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 * Link-time stand-ins for the eight htslib entry points Mapping.cpp references
 * (reference src/Mapping.cpp:43,612-620,656,679,734). The SAM text path of the
 * reference never calls them; requesting -bo with this build aborts loudly.
 * This lets the unmodified reference sources link without running htslib's
 * own build system. */
#include <stdio.h>
#include <stdlib.h>
static void die(const char *fn) { fprintf(stderr, "[oracle/_ref] %s: BAM output is not available in the oracle build\n", fn); abort(); }
void *hts_open_format(const char *fn, const char *mode, const void *fmt) { (void)fn; (void)mode; (void)fmt; die("hts_open_format"); return 0; }
int hts_close(void *fp) { (void)fp; die("hts_close"); return -1; }
void *sam_hdr_parse(int l, const char *t) { (void)l; (void)t; die("sam_hdr_parse"); return 0; }
int sam_hdr_write(void *fp, const void *h) { (void)fp; (void)h; die("sam_hdr_write"); return -1; }
int sam_parse1(void *s, void *h, void *b) { (void)s; (void)h; (void)b; die("sam_parse1"); return -1; }
int sam_write1(void *fp, const void *h, const void *b) { (void)fp; (void)h; (void)b; die("sam_write1"); return -1; }
void *bam_init1(void) { die("bam_init1"); return 0; }
void bam_destroy1(void *b) { (void)b; die("bam_destroy1"); }
