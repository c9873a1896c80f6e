"""GPU probe of the other configs' shapes (C4: SE 100 bp @ 8 %, C5: -pacbio 7 kbp @ 15 %) on the available indexes: device time per stage,
work counters, reads/s (resident batch) and e2e (host buffers), with the reference timed on a sample. One JSON line per (index, mode).
Usage: python scripts/gpu_modes.py [--se N] [--pb N] [--prefixes a,b]"""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_util as pu
from kart_b200 import KartIndex, Mapper, synth

ap = argparse.ArgumentParser()
ap.add_argument("--se", type=int, default=2000000)
ap.add_argument("--pb", type=int, default=20000)
ap.add_argument("--pblen", type=int, default=7000)
ap.add_argument("--prefixes", default=",".join([pu.ECOLI_PREFIX, os.path.join(ROOT, "data", "_gen", "syn", "syn100")]))
ap.add_argument("--ref-se", type=int, default=200000)
ap.add_argument("--ref-pb", type=int, default=2000)
ap.add_argument("--check", type=int, default=200, help="reads compared with the oracle per mode")
ap.add_argument("--modes", default="se100,pacbio")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ncores = os.cpu_count() or 1
for prefix in a.prefixes.split(","):
    if not os.path.exists(prefix + ".bwt"):
        continue
    idx = KartIndex(prefix); g = pu.pac_genome(idx) if idx.l_pac > 500_000_000 else pu.genome_of(idx)
    m = Mapper(); m.upload_index(idx, expand_sa=True)
    for mode, n, L, err, kw, nref in (("se100", a.se, 100, 0.08, {}, a.ref_se), ("pacbio", a.pb, a.pblen, 0.15, dict(indel=0.01), a.ref_pb)):
        if n <= 0 or mode not in a.modes.split(","):
            continue
        r, _, pos = synth.simulate(g, n, L, err, seed=3, paired=False, **kw)
        flat, off = Mapper.pack_reads(r)
        m.set_params(pacbio=(mode == "pacbio"), paired=False)
        m.stage(flat, off)
        m.run()
        t = time.perf_counter(); reps = a.reps
        st = {}
        for _ in range(reps):
            m.run()
            for k, v in m.stage_ms().items():
                st[k] = st.get(k, 0.0) + v / reps
        dev_s = (time.perf_counter() - t) / reps
        w = m.work()
        t = time.perf_counter(); aln, _, cig = m.map_chunk(flat, off); e2e_s = time.perf_counter() - t
        t = time.perf_counter(); aln, _, cig = m.map_chunk(flat, off); e2e_s = min(e2e_s, time.perf_counter() - t)
        # parity spot check against the oracle
        bad = -1
        if a.check:
            m2 = pu.make_mapper(idx, pacbio=(mode == "pacbio"), paired=False)
            bad = pu.compare_singles(m2, pu.Oracle(prefix, pacbio=(mode == "pacbio")), r[:a.check if mode != "pacbio" else max(4, a.check // 20)])
            del m2
        out = {"index": os.path.basename(prefix), "mode": mode, "reads": n, "read_len": L, "error": err, "device_reads_per_s": n / dev_s, "device_ms": dev_s * 1e3,
               "e2e_reads_per_s": n / e2e_s, "bases_per_s_device": n * L / dev_s, "stage_ms": st, "work": w, "mapped_fraction": float((aln["score"] > 0).mean()),
               "nw_gcups": w["nw_cells"] / max(st.get("align", 0), 1e-9) / 1e6, "oracle_mismatches": bad}
        if os.path.exists(pu.REF_KART) and nref > 0:
            tmp = tempfile.mkdtemp(prefix="kartmodes")
            f = os.path.join(tmp, "r.fq"); synth.write_fastq(f, r[:nref], pos[:nref], 1, err)
            e = os.path.join(tmp, "e.fq"); synth.write_fastq(e, r[:1], pos[:1], 1, err)
            def run(x):
                t0 = time.perf_counter()
                subprocess.run([pu.REF_KART, "-silent", "-t", str(ncores), "-i", prefix, "-f", x, "-o", os.path.join(tmp, "o.sam")] + (["-pacbio"] if mode == "pacbio" else []), check=True, stdout=subprocess.DEVNULL)
                return time.perf_counter() - t0
            load = min(run(e), run(e)); tot = run(f)
            out["reference_reads_per_s"] = nref / max(tot - load, 1e-6); out["reference_cores"] = ncores; out["reference_sample_reads"] = nref
            subprocess.run(["rm", "-rf", tmp])
        print(json.dumps(out), flush=True)
    m.close()
