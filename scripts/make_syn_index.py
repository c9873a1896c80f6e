"""Builds a synthetic multi-contig genome (i.i.d. bases + injected repeat families, SURVEY.md Appendix B pilot, scaled) and
indexes it with the reference's own bwt_index (oracle/_ref), or with `kart index` of this repo (same bytes) when that is absent
or KART_INDEX_BUILDER=ours. Usage: python scripts/make_syn_index.py <Mbp> [contigs] [seed] [outdir]
Output: data/_gen/syn/syn<Mbp>.{fa,bwt,sa,pac,ann,amb} (git-ignored; travels to the GPU box with the snapshot)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kart_b200 import synth
mbp = int(sys.argv[1]); contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 4; seed = int(sys.argv[3]) if len(sys.argv) > 3 else 12345
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "data", "_gen", "syn"); os.makedirs(out, exist_ok=True)
prefix = os.path.join(out, "syn%d" % mbp)
t = time.time()
scale = mbp / 100.0
names, seqs = synth.make_genome(mbp * 1000000, contigs, seed, repeats=((3000, int(2000 * scale), 0.02), (300, int(20000 * scale), 0.05)))
synth.write_fasta(prefix + ".fa", names, seqs)
print("genome written %.1fs" % (time.time() - t), flush=True)
t = time.time()
ref_tool = os.path.join(ROOT, "oracle", "_ref", "bwt_index")
if os.path.exists(ref_tool) and os.environ.get("KART_INDEX_BUILDER", "reference") != "ours":
    subprocess.run([ref_tool, prefix + ".fa", prefix], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
else:   # our own builder writes the same bytes (tests/test_index_build.py), an order of magnitude faster
    subprocess.run([os.path.join(ROOT, "kart_b200", "bin", "kart"), "index", prefix + ".fa", prefix], check=True, stdout=subprocess.DEVNULL)
print("index built %.1fs" % (time.time() - t), flush=True)
os.remove(prefix + ".fa")
