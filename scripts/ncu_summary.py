#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) and an ncu launch list into small tracked files under profiles/.
Usage: python scripts/ncu_summary.py <tag> [pairs per launch] [index name]   (reads gpurun_out/<tag>_prof.ncu-rep, gpurun_out/<tag>_launches.csv)
Run it on the box that took the capture (or before touching the CUDA sources): the traffic file records the hash of kart_b200/csrc, and
bench.py only quotes a capture of the build it is running."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
        "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    tag = sys.argv[1]
    rep = os.path.join(ROOT, "gpurun_out", tag + "_prof.ncu-rep")
    out = os.path.join(ROOT, "profiles", tag + "_ncu_full_summary.md")
    lines = ["# ncu --set full --clock-control none summary (%s)" % tag, "", "One row group per captured launch; values straight from `ncu -i %s_prof.ncu-rep --page raw --csv`." % tag, ""]
    if os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            lines.append("## %s" % r[ki].split("(")[0])
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append("| %s | %s | %s |" % (w, r[i], units[i]))
            lines.append("")
        open(out, "w").write("\n".join(lines))
        print("wrote", out)
        # per-kernel DRAM traffic of the largest launch (the resident-batch launch), for bench.py's roofline.traffic
        import json
        tr = {}
        gi, ri, wi, ti = hdr.index("launch__grid_size"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[ki].split("(")[0]
            g = int(float(r[gi]))
            if name not in tr or g > tr[name]["grid"]:
                tr[name] = {"grid": g, "block": int(float(r[hdr.index("launch__block_size")])), "dram_bytes": float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]], "duration_ms": float(r[ti]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[ti], 1.0)}
        pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
        import hashlib
        h = hashlib.sha256()
        d = os.path.join(ROOT, "kart_b200", "csrc")
        for f in sorted(os.listdir(d)):
            h.update(open(os.path.join(d, f), "rb").read())
        json.dump({"tag": tag, "pairs_per_launch": pairs, "workload": sys.argv[3] if len(sys.argv) > 3 else "EcoliIdx", "source_sha": h.hexdigest()[:16], "kernels": tr}, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
    lc = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc, errors="replace")) if len(r) > 10 and r[0].isdigit()]
        agg = {}
        for r in rows:
            name = r[4].split("(")[0]
            a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
            a[0] += 1
            a[1] += float(r[-1]) / 1e3
        UPLOAD = ("k_reblock", "k_expand_sa", "k_build_ktab", "k_ref64")   # index upload, once per context: not part of a step
        tot = sum(v[1] for k, v in agg.items() if k not in UPLOAD)
        o = ["# ncu launch list (%s): gpu__time_duration.sum per kernel, --clock-control none" % tag, "",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --pairs 500000` (per-launch times are cold-cache and serialised; shares are what matter).", "",
             "| kernel | launches | total us | mean us | share of step (excl. upload kernels) | grid | block |", "|---|---|---|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = "" if k in UPLOAD else "%.1f %%" % (100 * v[1] / tot)
            o.append("| %s | %d | %.1f | %.1f | %s | %s | %s |" % (k, v[0], v[1], v[1] / v[0], share, v[2], v[3]))
        open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w").write("\n".join(o) + "\n")
        import shutil
        shutil.copy(lc, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
        print("wrote launches")
    for suffix in ("_bench.json", "_bench_ref.json", "_pytest.txt", "_gpu.txt"):
        p = os.path.join(ROOT, "gpurun_out", tag + suffix)
        if os.path.exists(p):
            import shutil
            shutil.copy(p, os.path.join(ROOT, "profiles", tag + suffix))


if __name__ == "__main__":
    main()
