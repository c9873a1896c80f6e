"""Whole-program comparison on the GPU box: kart_b200/bin/kart against the unmodified reference (oracle/_ref/kart) on the
same FASTQ files. Prints one JSON line with both wall times, reads/s net of index load, and whether the SAM files are
byte-identical (raw vs `-t 1` when --t1 is given, and after `LC_ALL=C sort` vs `-t <nproc>`).
Usage: python scripts/cli_compare.py [--pairs N] [--prefix P] [--error E] [--mode pe|se|pacbio] [--t1] [--len L] [--extra "flags"]"""
import argparse, hashlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu
from kart_b200 import KartIndex, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1000000)
ap.add_argument("--prefix", default=None)
ap.add_argument("--error", type=float, default=0.02)
ap.add_argument("--mode", default="pe", choices=["pe", "se", "pacbio"])
ap.add_argument("--len", type=int, default=0)
ap.add_argument("--t1", action="store_true")
ap.add_argument("--ours-only", action="store_true")
ap.add_argument("--no-compare", action="store_true", help="time both programs, skip sorting and comparing the outputs (large inputs)")
ap.add_argument("--extra", default="")
ap.add_argument("--diff-out", default=None, help="write the first differing lines (ours vs reference) to this file")
ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
a = ap.parse_args()
prefix = a.prefix or pu.default_prefix()
idx = KartIndex(prefix); genome = pu.pac_genome(idx) if idx.l_pac > 500_000_000 else pu.genome_of(idx)
L = a.len or {"pe": 150, "se": 100, "pacbio": 7000}[a.mode]
tmp = tempfile.mkdtemp(prefix="kartcli")
t = time.time()
if a.mode == "pe":
    f1, f2 = synth.make_reads(genome, os.path.join(tmp, "r"), a.pairs, L, a.error, seed=1)
    files = ["-f", f1, "-f2", f2]; n_reads = 2 * a.pairs
else:
    kw = dict(indel=0.01) if a.mode == "pacbio" else {}
    f1, _ = synth.make_reads(genome, os.path.join(tmp, "r"), a.pairs, L, a.error, seed=3, paired=False, **kw)
    files = ["-f", f1] + (["-pacbio"] if a.mode == "pacbio" else []); n_reads = a.pairs
e1 = os.path.join(tmp, "e_1.fq")
open(e1, "wb").write(b"".join(open(f1, "rb").readlines()[:4]))
empty = ["-f", e1] + (["-pacbio"] if a.mode == "pacbio" else [])
gen_s = time.time() - t
extra = a.extra.split()
OURS = os.environ.get("KART_BIN", os.path.join(ROOT, "kart_b200", "bin", "kart"))


def run(binary, threads, args, out):
    t = time.perf_counter()
    subprocess.run([binary, "-silent", "-t", str(threads), "-i", prefix] + args + ["-o", out] + (extra if binary == OURS else []), check=True, stdout=subprocess.DEVNULL)
    return time.perf_counter() - t


def md5(path, sort=False):
    if sort:
        s = path + ".sorted"
        subprocess.run("LC_ALL=C sort -S 4G --parallel=8 %s > %s" % (path, s), shell=True, check=True)
        path = s
    h = hashlib.md5()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


res = {"prefix": os.path.basename(prefix), "mode": a.mode, "reads": n_reads, "read_len": L, "error": a.error, "threads": a.threads, "gen_s": round(gen_s, 2)}
ours_sam, ref_sam, ref1_sam = (os.path.join(tmp, x) for x in ("ours.sam", "ref.sam", "ref1.sam"))
res["ours_load_s"] = min(run(OURS, a.threads, empty, os.path.join(tmp, "e.sam")) for _ in range(2))
res["ours_total_s"] = min(run(OURS, a.threads, files, ours_sam) for _ in range(2))
if a.ours_only:
    res["ours_reads_per_s"] = n_reads / max(res["ours_total_s"] - res["ours_load_s"], 1e-6)
    print(json.dumps(res)); subprocess.run(["rm", "-rf", tmp]); sys.exit(0)
res["ref_load_s"] = min(run(pu.REF_KART, a.threads, empty, os.path.join(tmp, "e.sam")) for _ in range(2))
res["ref_total_s"] = run(pu.REF_KART, a.threads, files, ref_sam)
# net of start-up; start-up (CUDA context creation) varies by a few tenths of a second between runs, so at this size the
# difference can vanish: report null then rather than a made-up rate
net = res["ours_total_s"] - res["ours_load_s"]
res["ours_reads_per_s"] = n_reads / net if net > 0.05 else None
res["ref_reads_per_s"] = n_reads / max(res["ref_total_s"] - res["ref_load_s"], 1e-6)
res["ours_reads_per_s_incl_load"] = n_reads / res["ours_total_s"]
res["ref_reads_per_s_incl_load"] = n_reads / res["ref_total_s"]
res["sam_bytes"] = os.path.getsize(ours_sam)
res["whole_program_ratio"] = res["ref_total_s"] / res["ours_total_s"]
if a.no_compare:
    res["sam_lines"] = [int(subprocess.run("grep -vc '^@' %s" % f, shell=True, capture_output=True, text=True).stdout.strip() or 0) for f in (ours_sam, ref_sam)]
    print(json.dumps(res)); subprocess.run(["rm", "-rf", tmp]); sys.exit(0)
res["sorted_identical_to_ref_tN"] = md5(ours_sam, True) == md5(ref_sam, True)


def dump_diff(a_path, b_path, tag):
    if not a.diff_out:
        return
    d = subprocess.run("diff %s %s | head -60" % (a_path, b_path), shell=True, capture_output=True, text=True).stdout
    n = subprocess.run("diff %s %s | grep -c '^<'" % (a_path, b_path), shell=True, capture_output=True, text=True).stdout.strip()
    with open(a.diff_out, "a") as fh:
        fh.write("== %s: %s differing lines\n%s\n" % (tag, n, d))


if not res["sorted_identical_to_ref_tN"]:
    dump_diff(ours_sam + ".sorted", ref_sam + ".sorted", "ours vs ref -t %d (sorted)" % a.threads)
if a.t1:
    res["ref_t1_total_s"] = run(pu.REF_KART, 1, files, ref1_sam)
    res["raw_identical_to_ref_t1"] = md5(ours_sam) == md5(ref1_sam)
    if not res["raw_identical_to_ref_t1"]:
        dump_diff(ours_sam, ref1_sam, "ours vs ref -t 1 (raw)")
    if not res["sorted_identical_to_ref_tN"]:
        res["ref_tN_sorted_identical_to_ref_t1"] = md5(ref_sam, True) == md5(ref1_sam, True)
print(json.dumps(res))
subprocess.run(["rm", "-rf", tmp])
