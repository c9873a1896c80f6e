"""Builds a synthetic multi-contig genome (i.i.d. bases + injected repeat families, SURVEY.md Appendix B pilot, scaled) and its index.
Builder, in this order unless KART_INDEX_BUILDER says otherwise:
  gpu        `kart index -gpu -pac` (csrc/kb_index_build.cu) from a .pac / .ann written straight from the generator: seconds at 3.1 Gbp
  ours       `kart index` of this repo on the host threads (same bytes, minutes)
  reference  the reference's own bwt_index (oracle/_ref; hours at 3.1 Gbp)
Usage: python scripts/make_syn_index.py <Mbp> [contigs] [seed] [outdir]
Output: data/_gen/syn/syn<Mbp>.{bwt,sa,pac,ann,amb} (git-ignored; small ones travel to the GPU box with the snapshot)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kart_b200 import synth
mbp = int(sys.argv[1]); contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 4; seed = int(sys.argv[3]) if len(sys.argv) > 3 else 12345
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "data", "_gen", "syn"); os.makedirs(out, exist_ok=True)
prefix = os.path.join(out, "syn%d" % mbp)
KART = os.path.join(ROOT, "kart_b200", "bin", "kart")
want = os.environ.get("KART_INDEX_BUILDER", "gpu")
scale = mbp / 100.0
t = time.time()
names, lens, codes = synth.make_genome_codes(mbp * 1000000, contigs, seed, repeats=((3000, int(2000 * scale), 0.02), (300, int(20000 * scale), 0.05)))
print("genome generated %.1fs" % (time.time() - t), flush=True)
done = False
if want == "gpu":
    t = time.time()
    synth.write_pac_ann(prefix, names, lens, codes)
    print("pac written %.1fs" % (time.time() - t), flush=True)
    t = time.time()
    r = subprocess.run([KART, "index", "-gpu", "-pac", prefix], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    done = r.returncode == 0 and os.path.exists(prefix + ".sa")
    print("index built on the GPU %.1fs" % (time.time() - t) if done else "GPU builder not usable here (%s): host builder instead" % r.stdout.strip()[-200:], flush=True)
if not done:
    avail_gb = 0
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            avail_gb = int(ln.split()[1]) >> 20
    if avail_gb < mbp * 23 // 1000 + 2:   # the host builder holds 2G x 8 B of suffix positions (4 B up to 2.1 Gbp) next to the text
        sys.exit("the host builder needs ~%d GB for %d Mbp, this host has %d GB available" % (mbp * 23 // 1000 + 2, mbp, avail_gb))
    t = time.time()
    cuts = [0]
    for ln in lens:
        cuts.append(cuts[-1] + int(ln))
    g = synth._ACGT[codes]
    synth.write_fasta(prefix + ".fa", names, [g[cuts[i]:cuts[i + 1]] for i in range(len(names))])
    del g
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "bwt_index")
    if want == "reference" and os.path.exists(ref_tool):
        subprocess.run([ref_tool, prefix + ".fa", prefix], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    else:   # our own host builder writes the same bytes (tests/test_index_build.py), an order of magnitude faster than the reference
        subprocess.run([KART, "index", prefix + ".fa", prefix], check=True, stdout=subprocess.DEVNULL)
    os.remove(prefix + ".fa")
    print("index built on the host %.1fs" % (time.time() - t), flush=True)
