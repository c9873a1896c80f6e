#!/usr/bin/env python
"""bench.py -- mapped reads/s of the kart_b200 hot path on BASELINE.json's C2 workload (E. coli, 2x150 bp @ 2 % error).

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


def ncu_traffic(kernel, reads, syn):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture of this kind of workload (C2, or a
    --prefix index: tags ending in 'syn'), scaled linearly from the captured launch's read count (one thread per read: grid x block)
    to this launch's."""
    import glob
    import re
    files = [f for f in glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")) if ("syn" in os.path.basename(f)) == bool(syn)]
    if not files:
        return None, None
    f = max(files, key=lambda p: int(re.findall(r"r(\d+)", os.path.basename(p))[0]))
    d = json.load(open(f))
    norm = {re.sub(r"^void |<.*$", "", name): v for name, v in d["kernels"].items()}   # "void k_fm_seed<10>" -> "k_fm_seed"
    k = norm.get(kernel) or norm.get(kernel + "_q")   # k_fm_seed_q: the lane-queue seeding kernel
    if not k:
        return None, None
    per_read = norm.get("k_segments") or k                # one thread per read there; k_fm_seed_q runs a fixed grid
    cap_reads = per_read["grid"] * per_read.get("block", 128)
    return k["dram_bytes"] * reads / cap_reads, "%s (captured at %d reads/launch, scaled by reads)" % (os.path.basename(f), cap_reads)


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


def workload(pairs, seed, prefix=None, err=0.02):
    import parity_util as pu
    from kart_b200 import KartIndex, synth
    prefix = prefix or pu.default_prefix()
    idx = KartIndex(prefix)
    genome = pu.genome_of(idx)
    r1, r2, pos = synth.simulate(genome, pairs, 150, err, seed=seed)
    return prefix, idx, r1, r2, pos


def time_reference(prefix, r1, r2, pos, sample_pairs, threads, tmp):
    """Unmodified reference binary on a bounded sample; index-load time (same command, 2 reads) is subtracted."""
    import parity_util as pu
    from kart_b200 import synth
    f1, f2 = os.path.join(tmp, "s_1.fq"), os.path.join(tmp, "s_2.fq")
    synth.write_fastq(f1, r1[:sample_pairs], pos[:sample_pairs], 1, 0.02)
    synth.write_fastq(f2, r2[:sample_pairs], pos[:sample_pairs], 2, 0.02)
    e1, e2 = os.path.join(tmp, "e_1.fq"), os.path.join(tmp, "e_2.fq")
    synth.write_fastq(e1, r1[:1], pos[:1], 1, 0.02)
    synth.write_fastq(e2, r2[:1], pos[:1], 2, 0.02)

    def run(a, b):
        t = time.perf_counter()
        subprocess.run([pu.REF_KART, "-silent", "-t", str(threads), "-i", prefix, "-f", a, "-f2", b, "-o", os.path.join(tmp, "ref.sam")],
                       check=True, stdout=subprocess.DEVNULL)
        return time.perf_counter() - t
    load = min(run(e1, e2) for _ in range(2))
    total = run(f1, f2)
    return 2 * sample_pairs / max(total - load, 1e-6), total, load


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kart_b200", choices=["kart_b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="read pairs per step per GPU (C2: 1M pairs)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=400_000)
    ap.add_argument("--full-sa", type=int, default=1, help="expand the sampled SA into a full SA in HBM at upload")
    ap.add_argument("--prefix", default=None, help="index prefix (default: the E. coli index of config C2); e.g. data/_gen/syn/syn400 for the HBM-bound regime")
    ap.add_argument("--error", type=float, default=0.02)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import parity_util as pu
    ncores = os.cpu_count() or 1
    wl = "C2: E. coli K-12 (4.64 Mbp, index from test/ecoli.fa)" if not args.prefix else "index %s" % os.path.basename(args.prefix)
    config = {"workload": "%s, %d synthetic paired-end reads 2x150 bp @ %g%% error per GPU per step, seed %d+rank" % (wl, 2 * args.pairs, 100 * args.error, 1),
              "reads_per_step_per_gpu": 2 * args.pairs, "read_len": 150, "error_rate": args.error, "full_sa_in_hbm": bool(args.full_sa),
              "l2": ("read batch (%.0f MB) exceeds L2; " % (2 * args.pairs * 150 / 1e6)) +
                    ("the E. coli FM-index (4.6 MB) is L2-resident by construction of this config" if not args.prefix else
                     "index %s is larger than L2 when its .bwt exceeds 126 MB" % os.path.basename(args.prefix))}

    if args.impl == "reference":
        if rank != 0:
            return
        if not os.path.exists(pu.REF_KART):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/kart was not built (needs /root/reference at build time)"}))
            return
        prefix, idx, r1, r2, pos = workload(args.cpu_sample_pairs, 1, args.prefix, args.error)
        tmp = tempfile.mkdtemp(prefix="kartbench")
        vals = []
        for it in range(args.warmup + args.steps):
            v, total, load = time_reference(prefix, r1, r2, pos, args.cpu_sample_pairs, ncores, tmp)
            if it >= args.warmup:
                vals.append((v, total - load))
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([t for _, t in vals]) * 1e3)
        sample = "%d reads (first %d pairs of the C2 stream), kart -t %d, index-load time subtracted" % (2 * args.cpu_sample_pairs, args.cpu_sample_pairs, ncores)
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
    prefix, idx, r1, r2, pos = workload(args.pairs, 1 + rank, args.prefix, args.error)
    reads = pu.interleave(r1, r2)
    n = reads.shape[0]
    m = Mapper(device=local)
    m.upload_index(idx, expand_sa=bool(args.full_sa))
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
    stage_sum = {k: 0.0 for k in STAGES + ["total"]}
    e0.record(stream)
    for _ in range(args.steps):
        m.run()
        for k, v in m.stage_ms().items():
            stage_sum[k] += v
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    work = m.work()
    # ---- end-to-end leg: pinned host buffers in, host results out ----
    for _ in range(max(1, args.warmup // 2)):
        aln, pairs, cig = m.map_chunk(flat, off, est, out=out_bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        aln, pairs, cig = m.map_chunk(flat, off, est, out=out_bufs)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = int(flat.nbytes + off.nbytes + (n // 2) * 4)
    d2h = int(aln.nbytes + cig.nbytes + pairs[:n // 2].nbytes)
    mapped = int((aln["score"] > 0).sum())
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_reads = n * world
    value = total_reads * args.steps / (dev_ms / 1e3)
    e2e = total_reads * args.steps / (e2e_ms / 1e3)
    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md "Kernels") ----
    per = {k: stage_sum[k] / args.steps for k in STAGES}
    dom = max(per, key=per.get)
    alg = {"fm_seed": 32.0 * work["occ_blocks"], "sa_locate": 32.0 * work["lf_steps"] + 8.0 * work["seeds"]}
    peak, which = peaks()
    rk = dom if dom in alg else "fm_seed"
    achieved = alg[rk] / (per[rk] / 1e3) / 1e9
    traffic, traffic_src = ncu_traffic("k_" + rk, n, args.prefix)
    roof = {"kernel": "k_" + rk, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": which, "dominant_kernel": "k_" + dom, "share_of_step": per[rk] / max(sum(per.values()), 1e-9),
            "note": ("E. coli index is L2-resident: achieved GB/s is L2->SM sector traffic expressed against the HBM copy peak; dram traffic is in profiles/"
                     if not args.prefix else "random 32-byte sector reads of the Occ blocks and seeding table")}
    rs = random_sector_peak()
    if rs:
        roof["random_sector_peak"] = rs
        if args.prefix:
            roof["frac_of_random_sector_peak"] = achieved / rs["independent_gbs"]
    if numa:
        config["host_binding"] = "each rank bound to its GPU's " + numa.split(" for ")[0]
    out = {"metric": "mapped reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/int32", "data": "synthetic",
           "config": config, "clocks": sampler.summary(),
           "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
           "gpu_launches": int(work["launches"]) * args.steps, "roofline": roof,
           "stage_ms": per, "mapped_fraction": mapped / n,
           "work_per_step": work, "seed_occ_gbs": alg["fm_seed"] / (per["fm_seed"] / 1e3) / 1e9,
           # work-equivalent rate: the traffic SURVEY 8(d) assigns to the reference's algorithm (64 B per extension step), which the
           # unique-tail seeding no longer moves
           "seed_ref_equiv_gbs": 64.0 * work["ext_steps"] / (per["fm_seed"] / 1e3) / 1e9,
           "nw_gcups": work["nw_cells"] / (per["align"] / 1e3) / 1e9}
    # ---- CPU baseline: the unmodified reference on this box's host cores, bounded sample ----
    if args.cpu_sample_pairs <= 0:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "skipped (--cpu-sample-pairs 0)"}
    elif os.path.exists(pu.REF_KART):
        tmp = tempfile.mkdtemp(prefix="kartbench")
        sp = min(args.cpu_sample_pairs, args.pairs)
        v, total, load = time_reference(prefix, r1, r2, pos, sp, ncores, tmp)
        out["cpu_baseline"] = {"value": v, "unit": "reads/s", "cores": ncores, "kind": "reference",
                               "sample": "%d reads (first %d pairs of the step's batch), oracle/_ref/kart -t %d, %.2f s map + %.2f s index load (subtracted)" % (2 * sp, sp, ncores, total - load, load)}
    else:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "oracle/_ref/kart not built"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
