"""CPU-side parity check on an index too large to ship to the GPU box (3.1 Gbp: 33-bit text positions, the 64-bit row kernel,
MinSeedLength 16): the device code compiled for the host (tests/emul, TEST INFRASTRUCTURE) behind the real CLI host code against
the unmodified reference `kart -t 1`, byte for byte, on C3/C4/C5-shaped reads.
Usage: python scripts/big_index_check.py <index prefix> [pairs=20000] [se=20000] [pacbio=200]   -> one JSON line per mode"""
import hashlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu
from kart_b200 import KartIndex, synth

prefix = sys.argv[1]
n_pe = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n_se = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
n_pb = int(sys.argv[4]) if len(sys.argv) > 4 else 200
EMUL = os.path.join(ROOT, "tests", "emul", "kart_emul")
t = time.time()
idx = KartIndex(prefix); genome = pu.genome_of(idx)
total = int(sum(len(c) for c in genome))
print("index loaded: %d contigs, %d bp, 2G = %d (%s 2^32) in %.0f s" % (len(genome), total, 2 * total, ">" if 2 * total > 2 ** 32 else "<", time.time() - t), file=sys.stderr, flush=True)
tmp = tempfile.mkdtemp(prefix="kartbig")
env = dict(os.environ, KB_KTAB_K=os.environ.get("KB_KTAB_K", "10"))   # the full 4^14 table is 8.6 GB and a CPU would take minutes to fill it


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def run(binary, args, out, e=None):
    t0 = time.time()
    subprocess.run([binary, "-silent", "-t", "1", "-i", prefix] + args + ["-o", out], check=True, stdout=subprocess.DEVNULL, env=e)
    return time.time() - t0


for mode, n, L, err, kw, flags in (("pe150@1%", n_pe, 150, 0.01, {}, []), ("se100@8%", n_se, 100, 0.08, dict(paired=False), []),
                                   ("pacbio7k@15%", n_pb, 7000, 0.15, dict(paired=False, indel=0.01), ["-pacbio"])):
    if n <= 0:
        continue
    f1, f2 = synth.make_reads(genome, os.path.join(tmp, mode.split("@")[0]), n, L, err, seed=5, **kw)
    files = (["-f", f1, "-f2", f2] if f2 else ["-f", f1]) + flags
    ours, ref = os.path.join(tmp, "ours.sam"), os.path.join(tmp, "ref.sam")
    t_ours = run(EMUL, files, ours, env)
    t_ref = run(pu.REF_KART, files, ref)
    same = md5(ours) == md5(ref)
    mapped = sum(1 for ln in open(ours) if not ln.startswith("@") and ln.split("\t")[2] != "*")
    hi = sum(1 for ln in open(ours) if not ln.startswith("@") and ln.split("\t")[2] != "*" and int(ln.split("\t")[3]) > 0 and ln.split("\t")[2] in ("chr%d" % k for k in range(18, 25)))
    res = {"index": os.path.basename(prefix), "two_g": 2 * total, "mode": mode, "reads": n * (2 if f2 else 1), "byte_identical_to_kart_t1": same,
           "mapped_records": mapped, "records_on_last_contigs(beyond 2^32 on the reverse strand / 2^31 forward)": hi, "emul_s": round(t_ours, 1), "ref_s": round(t_ref, 1)}
    if not same:
        res["first_diff"] = subprocess.run("diff %s %s | head -6" % (ours, ref), shell=True, capture_output=True, text=True).stdout[:1500]
    print(json.dumps(res), flush=True)
subprocess.run(["rm", "-rf", tmp])
