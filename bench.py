#!/usr/bin/env python
"""bench.py -- mapped reads/s of the kart_b200 hot path on BASELINE.json's C3 workload: synthetic 3.1 Gbp reference (random + injected
repeats, 24 contigs), wgsim-model paired-end reads 2x150 bp @ 1 % error, 10 M pairs sharded over 8 GPUs = 1.25 M pairs per GPU per
step (weak scaling: every GPU maps one shard). The index (5.4 GB of .bwt/.sa/.pac) cannot travel in a snapshot, so the first run on
a box builds it with this repo's `kart index` from a seeded genome and caches it under data/_gen/syn/ (--workload c2 selects the
E. coli config instead; a host without the memory for the 3.1 Gbp build falls back to a smaller genome and says so).

One "step" = one pass of the whole hot path (fm_seed -> sa_locate -> cand_pair -> rescue -> report -> finalize) over one batch
of synthetic paired reads. Prints ONE JSON line (see the task contract): `value` is device-timed with the batch resident in
HBM, `e2e` goes through kb_map_chunk with pinned HOST buffers (H2D + kernels + D2H inside the timed region), `roofline`
describes the dominant kernel, `cpu_baseline` is the unmodified reference (oracle/_ref/kart) on the box's host cores.
`--impl reference` times that CPU reference instead (same metric/config), rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STAGES = ["fm_seed", "sa_locate", "cand_pair", "rescue", "segments", "align", "assemble", "finalize"]


def random_sector_peak():
    """Measured ceiling of random 32-byte sector reads on this GPU (scripts/gpu_randsector.cu, committed result): what an
    FM-index walk over an index larger than L2 can reach at best; the HBM copy peak is not reachable with 32-byte requests."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*randsector*.json*"))):
        for ln in open(f):
            try:
                d = json.loads(ln)
            except ValueError:
                continue
            if d.get("buffer_mib", 0) >= 4096 and (best is None or d["buffer_mib"] < best[0]["buffer_mib"]):
                best = (d, os.path.basename(f))
    if not best:
        return None
    return {"independent_gbs": best[0]["independent_gbs"], "dependent_chain_gbs": best[0]["dependent_1280_per_sm_gbs"], "buffer_mib": best[0]["buffer_mib"], "source": "profiles/" + best[1]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons}


C3_MBP, C3_CONTIGS, C3_SEED = 3100, 24, 12345


def ensure_index(args, rank):
    """Index prefix and workload description. C3: rank 0 builds the synthetic genome's index once per box (scripts/make_syn_index.py
    with this repo's `kart index -gpu`: files byte-identical to the reference builder's, seconds instead of hours; the host builder when
    the device builder cannot run), the other ranks wait."""
    import parity_util as pu
    if args.prefix:
        return args.prefix, "index %s" % os.path.basename(args.prefix)
    if args.workload == "c2":
        return pu.default_prefix(), "C2: E. coli K-12 (4.64 Mbp, index from test/ecoli.fa)"
    mbp = C3_MBP
    avail_gb = 0
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            avail_gb = int(ln.split()[1]) >> 20
    prefix = os.path.join(ROOT, "data", "_gen", "syn", "syn%d" % mbp)
    if not os.path.exists(prefix + ".ok") and avail_gb < mbp * 6 // 1000 + 4:
        mbp = 100   # the generator, the .pac and what the GPU builder hands back need ~6 GB per Gbp of host memory
        prefix = os.path.join(ROOT, "data", "_gen", "syn", "syn%d" % mbp)
    if not os.path.exists(prefix + ".ok"):
        if rank == 0:
            t = time.time()
            if not (os.path.exists(prefix + ".bwt") and os.path.exists(prefix + ".sa") and os.path.exists(prefix + ".pac") and os.path.exists(prefix + ".ann")):
                subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_syn_index.py"), str(mbp), str(C3_CONTIGS if mbp >= 1000 else 4), str(C3_SEED)],
                               check=True, stdout=sys.stderr)
            open(prefix + ".ok", "w").write("built in %.0f s\n" % (time.time() - t))
        else:
            while not os.path.exists(prefix + ".ok"):
                time.sleep(2)
    what = "C3: synthetic %.1f Gbp reference (i.i.d. bases + injected 3-kbp and 300-bp repeat families, %d contigs, seed %d)" % (mbp / 1000.0, C3_CONTIGS if mbp >= 1000 else 4, C3_SEED)
    if mbp != C3_MBP:
        what += " -- REDUCED from 3.1 Gbp: this host has %d GB of memory available for the index build" % avail_gb
    return prefix, what


def workload(pairs, seed, prefix, err):
    import parity_util as pu
    from kart_b200 import KartIndex, synth
    idx = KartIndex(prefix)
    r1, r2, pos = synth.simulate(pu.pac_genome(idx), pairs, 150, err, seed=seed)
    return idx, r1, r2, pos


class ReferenceRunner:
    """The unmodified reference binary (oracle/_ref/kart -t <cores>) on FASTQ files of a bounded sample of the step's reads. Its
    index-load time (the same command on one pair) is measured once and subtracted from every run."""

    def __init__(self, prefix, r1, r2, pos, sample_pairs, threads, err):
        import parity_util as pu
        from kart_b200 import synth
        self.pu, self.prefix, self.threads, self.n = pu, prefix, threads, sample_pairs
        self.tmp = tempfile.mkdtemp(prefix="kartbench")
        self.f = [os.path.join(self.tmp, x) for x in ("s_1.fq", "s_2.fq", "e_1.fq", "e_2.fq")]
        synth.write_fastq(self.f[0], r1[:sample_pairs], pos[:sample_pairs], 1, err)
        synth.write_fastq(self.f[1], r2[:sample_pairs], pos[:sample_pairs], 2, err)
        synth.write_fastq(self.f[2], r1[:1], pos[:1], 1, err)
        synth.write_fastq(self.f[3], r2[:1], pos[:1], 2, err)
        self.load = min(self._run(self.f[2], self.f[3]) for _ in range(2))

    def _run(self, a, b):
        t = time.perf_counter()
        subprocess.run([self.pu.REF_KART, "-silent", "-t", str(self.threads), "-i", self.prefix, "-f", a, "-f2", b, "-o", os.path.join(self.tmp, "ref.sam")],
                       check=True, stdout=subprocess.DEVNULL)
        return time.perf_counter() - t

    def step(self):
        """(reads/s net of index load, seconds of mapping)"""
        total = self._run(self.f[0], self.f[1])
        return 2 * self.n / max(total - self.load, 1e-6), total - self.load

    def close(self):
        subprocess.run(["rm", "-rf", self.tmp])


def bind_to_gpu_numa_node(local):
    """Pins this rank to the cores of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated (first touch
    places them there): with 8 ranks each moving 0.46 GB per step over PCIe, remote-node buffers halve the end-to-end rate.
    Best effort: returns a short description, or None when the topology cannot be read."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus) for %s" % (node, len(cpus), bdf)
    except Exception:
        return None


def source_sha():
    """Hash of the CUDA sources: an ncu traffic capture is only quoted for the build it was taken from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "kart_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel, reads, workload_key):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of THIS workload and THIS build
    (profiles/*_traffic.json written by scripts/ncu_summary.py on the GPU box, with the hash of the CUDA sources and the workload it
    profiled), scaled by reads per launch. A capture of another build or workload is not quoted: traffic is then null."""
    import glob
    import re
    best = None
    sha = source_sha()
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):   # file times do not survive a checkout: the capture of this build wins, else the last tag
        d = json.load(open(f))
        if d.get("workload") != workload_key:
            continue
        if best is None or best[1].get("source_sha") != sha:
            best = (f, d)
    if best is None:
        return None, "no ncu capture of workload %s under profiles/" % workload_key
    f, d = best
    if d.get("source_sha") != source_sha():
        return None, "%s was captured from another build (sources changed since): not quoted" % os.path.basename(f)
    norm = {re.sub(r"^void |<.*$", "", name): v for name, v in d["kernels"].items()}
    k = norm.get(kernel) or norm.get(kernel + "_q")
    if not k:
        return None, "%s has no launch of %s" % (os.path.basename(f), kernel)
    cap_reads = 2 * d["pairs_per_launch"]
    return k["dram_bytes"] * reads / cap_reads, "%s (same build, captured at %d reads/launch, scaled by reads)" % (os.path.basename(f), cap_reads)


def program_leg(prefix, r1, r2, pos, pairs, err, threads):
    """Whole program, FASTQ in -> SAM out, same files: kart_b200/bin/kart against oracle/_ref/kart -t <cores>. What a user of the
    drop-in sees; start-up (index load, CUDA context) is inside both wall times and also reported on its own."""
    import parity_util as pu
    from kart_b200 import synth
    ours = os.path.join(ROOT, "kart_b200", "bin", "kart")
    if not (os.path.exists(ours) and os.path.exists(pu.REF_KART)):
        return None
    tmp = tempfile.mkdtemp(prefix="kartprog")
    f = [os.path.join(tmp, x) for x in ("p_1.fq", "p_2.fq", "e_1.fq", "e_2.fq")]
    synth.write_fastq(f[0], r1[:pairs], pos[:pairs], 1, err); synth.write_fastq(f[1], r2[:pairs], pos[:pairs], 2, err)
    synth.write_fastq(f[2], r1[:1], pos[:1], 1, err); synth.write_fastq(f[3], r2[:1], pos[:1], 2, err)

    def run(binary, a, b, out):
        t = time.perf_counter()
        subprocess.run([binary, "-silent", "-t", str(threads), "-i", prefix, "-f", a, "-f2", b, "-o", os.path.join(tmp, out)], check=True, stdout=subprocess.DEVNULL)
        return time.perf_counter() - t
    res = {"reads": 2 * pairs, "threads": threads}
    res["ours_startup_s"] = run(ours, f[2], f[3], "e.sam")
    # CUDA context creation on a box whose GPU is held by another process (this one) takes anything between 0.3 and 3+ s (r19-r36 traces):
    # ours is run twice and the faster run counts, both are reported; the reference (tens of seconds) runs once
    res["ours_runs_s"] = [run(ours, f[0], f[1], "ours.sam") for _ in range(2)]
    res["ours_total_s"] = min(res["ours_runs_s"])
    res["ref_startup_s"] = run(pu.REF_KART, f[2], f[3], "e.sam"); res["ref_total_s"] = run(pu.REF_KART, f[0], f[1], "ref.sam")
    # `kart -t N` writes its chunks in completion order and its threads race on EstDistance (its own -t 1 output differs from its
    # -t 16 output in a handful of pairs per million, profiles/r21_cli_diff.txt); ours equals -t 1 byte for byte (tests). So: sorted
    # lines, and the number that differ, not a checksum.
    for x in ("ours", "ref"):
        subprocess.run("LC_ALL=C sort -S 2G --parallel=8 -o %s %s" % (os.path.join(tmp, x + ".sorted"), os.path.join(tmp, x + ".sam")), shell=True, check=True)
    d = subprocess.run("LC_ALL=C comm -3 %s %s | wc -l" % (os.path.join(tmp, "ours.sorted"), os.path.join(tmp, "ref.sorted")), shell=True, capture_output=True, text=True).stdout.split()
    res["sam_lines_differing_from_ref_tN"] = int(d[0]) if d else None
    res["sorted_sam_identical"] = res["sam_lines_differing_from_ref_tN"] == 0
    res["ours_reads_per_s"] = 2 * pairs / res["ours_total_s"]; res["ref_reads_per_s"] = 2 * pairs / res["ref_total_s"]
    net_o, net_r = res["ours_total_s"] - res["ours_startup_s"], res["ref_total_s"] - res["ref_startup_s"]
    res["ours_reads_per_s_net_of_startup"] = 2 * pairs / net_o if net_o > 0.05 else None
    res["ref_reads_per_s_net_of_startup"] = 2 * pairs / net_r if net_r > 0.05 else None
    subprocess.run(["rm", "-rf", tmp])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kart_b200", choices=["kart_b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"], help="BASELINE.json config: c3 = synthetic 3.1 Gbp reference, PE 2x150 @ 1 % (default); c2 = E. coli, PE 2x150 @ 2 %")
    ap.add_argument("--pairs", type=int, default=0, help="read pairs per step per GPU (default: c3 1.25 M = one of 8 shards of 10 M pairs; c2 1 M)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=1_250_000, help="pairs of the step's batch the CPU reference is timed on (default: the whole step, ~10 s on 16 cores; 0: skip)")
    ap.add_argument("--program-pairs", type=int, default=1_250_000, help="pairs for the whole-program leg (FASTQ in -> SAM out, both programs), N = 1 only; 0: skip")
    ap.add_argument("--in-flight", type=int, default=3, help="chunks in flight in the end-to-end leg (2 or 3)")
    ap.add_argument("--full-sa", type=int, default=1, help="expand the sampled SA into a full SA in HBM at upload")
    ap.add_argument("--prefix", default=None, help="map against this index instead of the workload's own")
    ap.add_argument("--error", type=float, default=-1.0)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import parity_util as pu
    ncores = os.cpu_count() or 1
    if args.pairs <= 0:
        args.pairs = 1_250_000 if args.workload == "c3" else 1_000_000
    if args.error < 0:
        args.error = 0.01 if args.workload == "c3" else 0.02
    if args.impl == "reference" and rank != 0:
        return
    prefix, wl = ensure_index(args, rank)
    big = os.path.getsize(prefix + ".bwt") > 126e6
    config = {"workload": "%s, %d synthetic paired-end reads 2x150 bp @ %g%% error per GPU per step (wgsim model, seed 1+rank)%s" %
                          (wl, 2 * args.pairs, 100 * args.error, "; one of 8 shards of C3's 10 M pairs per GPU" if args.workload == "c3" and not args.prefix and args.pairs == 1_250_000 else ""),
              "index": os.path.basename(prefix), "reads_per_step_per_gpu": 2 * args.pairs, "read_len": 150, "error_rate": args.error, "full_sa_in_hbm": bool(args.full_sa),
              "l2": ("read batch (%.0f MB) exceeds L2; " % (2 * args.pairs * 150 / 1e6)) +
                    ("the FM-index (%.1f GB of Occ blocks + seeding table + full SA) is far larger than the 126 MB L2" % (os.path.getsize(prefix + ".bwt") / 1e9) if big
                     else "the E. coli FM-index (4.6 MB) is L2-resident by construction of this config")}
    workload_key = os.path.basename(prefix)

    if args.impl == "reference":
        if not os.path.exists(pu.REF_KART):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/kart was not built (needs /root/reference at build time)"}))
            return
        sp = min(args.cpu_sample_pairs if args.cpu_sample_pairs > 0 else args.pairs, args.pairs)
        idx, r1, r2, pos = workload(sp, 1, prefix, args.error)
        ref = ReferenceRunner(prefix, r1, r2, pos, sp, ncores, args.error)
        vals = [ref.step() for _ in range(args.warmup + args.steps)][args.warmup:]
        ref.close()
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([t for _, t in vals]) * 1e3)
        sample = "each step = %d reads (%s) through oracle/_ref/kart -t %d, FASTQ in / SAM out; index load (%.1f s, measured once) subtracted" % (
            2 * sp, "the whole step's batch" if sp == args.pairs else "the first %d pairs of the step's batch" % sp, ncores, ref.load)
        print(json.dumps({"impl": "reference", "metric": "mapped reads/s", "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/int32", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": value, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": sample},
                          "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    from kart_b200 import Mapper
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; kart_b200 has no CPU path")
    # The whole-program leg (N = 1) runs BEFORE this process touches the GPU: a second CUDA context on the device makes the CLI's own
    # context creation and first allocations several times slower (r40 trace: kb_init 1.8-2.5 s and a 0.5-1.9 s first batch next to this
    # process's context, against 0.3-0.5 s and 0.2 s alone), which is not what a user of the program sees.
    early = None
    if world == 1 and args.program_pairs > 0:
        early = workload(args.pairs, 1 + rank, prefix, args.error)
        try:
            e2e_program = program_leg(prefix, early[1], early[2], early[3], min(args.program_pairs, args.pairs), args.error, ncores)
        except Exception as e:   # the headline numbers stand on their own
            e2e_program = {"error": str(e)[:200]}
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    json_fd = 1
    if world > 1:
        # NCCL writes its version banner to stdout; rank 0 must print ONE JSON line there: libraries get stderr as their stdout,
        # the JSON line goes to the saved descriptor
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idx, r1, r2, pos = early if early is not None else workload(args.pairs, 1 + rank, prefix, args.error)
    reads = pu.interleave(r1, r2)
    n = reads.shape[0]
    m = Mapper(device=local)
    t_up = time.perf_counter()
    m.upload_index(idx, expand_sa=bool(args.full_sa))
    upload_s = time.perf_counter() - t_up
    m.set_params(paired=True)
    # pinned host buffers for the end-to-end leg
    seq_pin = torch.empty(reads.size, dtype=torch.uint8).pin_memory()
    seq_pin.numpy()[:] = reads.reshape(-1)
    off_pin = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    off_pin.numpy()[:] = np.arange(n + 1, dtype=np.int64) * 150
    flat, off = seq_pin.numpy(), off_pin.numpy().view(np.uint64)
    est = np.full(n // 2, 1500, dtype=np.int32)
    # caller-owned pinned result buffers (kb_results_t)
    from kart_b200.binding import ALN_DTYPE, PAIR_DTYPE
    aln_pin = torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    pair_pin = torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    cig_pin = torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory()
    out_bufs = (aln_pin.numpy().view(ALN_DTYPE), pair_pin.numpy().view(PAIR_DTYPE), cig_pin.numpy().view(np.uint32))
    stream = torch.cuda.ExternalStream(m.lib.kb_cuda_stream(m.h), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: kernels only ----
    m.stage(flat, off, est)
    for _ in range(args.warmup):
        m.run()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_sum = {k: 0.0 for k in STAGES + ["total", "nw"]}
    e0.record(stream)
    for _ in range(args.steps):
        m.run()
        for k, v in m.stage_ms().items():
            stage_sum[k] += v
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    work = m.work()
    # ---- end-to-end leg: pinned host buffers in, host results out, through kb_map_chunk_packed (the 2-bit form a host that packs
    # while parsing hands over: 46 B per read over PCIe); then the same through kb_map_chunk on the text (160 B per read) ----
    code_pin = torch.empty(int(m.lib.kb_packed_words(__import__("ctypes").byref(m._reads_struct(flat, off)))), dtype=torch.int64).pin_memory()
    t0 = time.perf_counter()
    pk = m.pack(flat, off, code=code_pin.numpy().view(np.uint64), threads=min(ncores, 16))
    pack_ms = (time.perf_counter() - t0) * 1e3
    # (a) chunks in flight (kb_map_chunk_begin_packed / kb_map_chunk_end): what a mapper streaming a read file does -- chunk k+1 is
    # handed over while chunk k is being mapped, so its H2D copy and chunk k-1's D2H copy run under chunk k's kernels. Every step still
    # copies its reads in from pinned memory and its records out; the timed region holds all K steps' copies.
    aln_pin2 = torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    pair_pin2 = torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    cig_pin2 = torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory()
    outs = [out_bufs, (aln_pin2.numpy().view(ALN_DTYPE), pair_pin2.numpy().view(PAIR_DTYPE), cig_pin2.numpy().view(np.uint32))]
    depth = max(2, min(3, args.in_flight))
    for _ in range(depth - 2):
        outs.append((torch.empty(n * ALN_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(ALN_DTYPE),
                     torch.empty((n // 2) * PAIR_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(PAIR_DTYPE),
                     torch.empty(4 * n + 1024, dtype=torch.int32).pin_memory().numpy().view(np.uint32)))

    def in_flight(steps):
        pending, res_ = [], None
        for k in range(steps):
            pending.append(m.map_chunk_begin(flat, off, est, out=outs[k % depth], packed=pk))
            if len(pending) == depth:
                res_ = m.map_chunk_end(pending.pop(0))
        while pending:
            res_ = m.map_chunk_end(pending.pop(0))
        return res_
    in_flight(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    aln, pairs, cig = in_flight(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    aln_packed = aln.copy()
    # (b) one synchronous call per step (kb_map_chunk_packed: the chunk is cut into sub-batches that overlap inside the call)
    for _ in range(max(1, args.warmup // 2)):
        m.map_chunk(flat, off, est, out=out_bufs, packed=pk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.map_chunk(flat, off, est, out=out_bufs, packed=pk)
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    m.map_chunk(flat, off, est, out=out_bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        aln, pairs, cig = m.map_chunk(flat, off, est, out=out_bufs)
    barrier()
    e2e_text_s = time.perf_counter() - t0
    same_records = bool(all(np.array_equal(aln[f], aln_packed[f]) for f in ("pos", "flag", "chr", "mapq", "score", "sub_score", "tlen", "cig_len")))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = int(pk[0].n_words * 8 + pk[0].n_exc * 8 + off.nbytes + (n // 2) * 4)
    h2d_text = int(flat.nbytes + off.nbytes + (n // 2) * 4)
    d2h = int(aln.nbytes + cig.nbytes + pairs[:n // 2].nbytes)
    mapped = int((aln["score"] > 0).sum())
    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_text_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_text_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_reads = n * world
    value = total_reads * args.steps / (dev_ms / 1e3)
    e2e = total_reads * args.steps / (e2e_ms / 1e3)
    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md "Kernels") ----
    per = {k: stage_sum[k] / args.steps for k in STAGES}
    nw_ms = stage_sum["nw"] / args.steps
    dom = max(per, key=per.get)
    alg = {"fm_seed": 32.0 * work["occ_blocks"], "sa_locate": 32.0 * work["lf_steps"] + 8.0 * work["seeds"]}
    peak, which = peaks()
    rk = dom if dom in alg else "fm_seed"
    achieved = alg[rk] / (per[rk] / 1e3) / 1e9
    traffic, traffic_src = ncu_traffic("k_" + rk, n, workload_key)
    roof = {"kernel": "k_" + rk, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": which, "dominant_kernel": "k_" + dom, "share_of_step": per[rk] / max(sum(per.values()), 1e-9),
            "algorithmic_bytes": "32 B x sectors the algorithm asks for (seeding-table entry, Occ blocks while the interval has > 1 row, SA entry, reference words): %.1f per read" % (work["occ_blocks"] / max(n, 1)),
            "note": ("random 32-byte sector reads over a %.1f GB index: the ceiling for that access pattern is random_sector_peak, not the copy peak" % (os.path.getsize(prefix + ".bwt") / 1e9) if big
                     else "E. coli index is L2-resident: achieved GB/s is L2->SM sector traffic expressed against the HBM copy peak")}
    rs = random_sector_peak()
    if rs:
        roof["random_sector_peak"] = rs
        if big:
            roof["frac_of_random_sector_peak"] = achieved / rs["independent_gbs"]
    if numa:
        config["host_binding"] = "each rank bound to its GPU's " + numa.split(" for ")[0]
    # nw_alignment as its own roofline: cell updates over the time of the solver kernels alone; the bound is the INT32 pipe.
    # One cell = 3 recurrences = ~15 integer instructions in k_nw_tile (SASS count); 148 SMs x 128 INT32 lanes x SM clock.
    clocks = sampler.summary()
    int_peak = 148 * 128 * (clocks.get("sm_max_mhz") or 1965) * 1e6
    nw = {"kernels": "k_nw_tile<0..5>, k_nw_warp (side by side)", "cells_per_step": work["nw_cells"], "calls_per_step": work["nw_calls"], "kernel_ms": nw_ms,
          "gcups": work["nw_cells"] / max(nw_ms, 1e-9) / 1e6, "bound": "int32 pipe", "int32_ops_per_cell": 15, "peak_gcups": int_peak / 15 / 1e9}
    nw["frac"] = nw["gcups"] / nw["peak_gcups"]
    out = {"metric": "mapped reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/int32", "data": "synthetic",
           "config": config, "clocks": clocks,
           "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                   "entry": "kb_map_chunk_begin_packed / kb_map_chunk_end, %d chunks in flight: pinned 2-bit words + exception list in, pinned kb_aln_t / cigar / pair statistics out, every step" % depth,
                   "host_pack_ms_outside_timed_region": pack_ms, "records_equal_text_entry": same_records},
           "e2e_sync": {"value": total_reads * args.steps / (e2e_sync_ms / 1e3), "unit": "reads/s", "ms_per_step": e2e_sync_ms / args.steps,
                        "entry": "kb_map_chunk_packed, one synchronous call per step (sub-batches overlap inside the call)"},
           "e2e_text": {"value": total_reads * args.steps / (e2e_text_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d_text, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_text_ms / args.steps, "entry": "kb_map_chunk, synchronous: pinned read characters in"},
           "gpu_launches": int(work["launches"]) * args.steps, "roofline": roof, "nw": nw,
           "stage_ms": per, "mapped_fraction": mapped / n, "index_upload_s": upload_s,
           "work_per_step": work, "seed_occ_gbs": alg["fm_seed"] / (per["fm_seed"] / 1e3) / 1e9,
           # work-equivalent rate: the traffic SURVEY 8(d) assigns to the reference's algorithm (64 B per extension step), which the
           # unique-tail seeding no longer moves -- a work rate, not a bandwidth
           "seed_ref_equiv_gbs": 64.0 * work["ext_steps"] / (per["fm_seed"] / 1e3) / 1e9,
           "nw_gcups": nw["gcups"], "align_stage_gcups": work["nw_cells"] / (per["align"] / 1e3) / 1e9}
    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample ----
    if args.cpu_sample_pairs <= 0:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "skipped (--cpu-sample-pairs 0)"}
    elif os.path.exists(pu.REF_KART):
        sp = min(args.cpu_sample_pairs, args.pairs)
        ref = ReferenceRunner(prefix, r1, r2, pos, sp, ncores, args.error)
        v, secs = ref.step()
        ref.close()
        out["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": ncores, "kind": "reference",
                               "sample": "%d reads (%s), oracle/_ref/kart -t %d, %.2f s map + %.2f s index load (subtracted)" % (
                                   2 * sp, "the whole step's batch" if sp == args.pairs else "first %d pairs of the step's batch" % sp, ncores, secs, ref.load)}
    else:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "oracle/_ref/kart not built"}
    if world == 1 and args.program_pairs > 0:
        out["e2e_program"] = e2e_program
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
