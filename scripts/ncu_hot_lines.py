#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
Usage: python scripts/ncu_hot_lines.py <rep> <kernel-regex> [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre, "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    path, out, total_s, total_i = "", [], 0, 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]
        elif len(r) > 8 and r[0].isdigit():
            samples = int(r[6]) if r[6].isdigit() else 0
            inst = int(r[7]) if r[7].isdigit() else 0
            tinst = int(r[8]) if r[8].isdigit() else 0
            out.append((samples, inst, tinst, path, int(r[0]), r[1].strip()))
            total_s += samples
            total_i += inst
    out.sort(reverse=True)
    print("kernel %s: %d samples, %d warp instructions" % (kre, total_s, total_i))
    print("%7s %6s %10s %6s  %s" % ("samples", "%", "warp-inst", "thr/in", "line"))
    for s, i, t, p, ln, code in out[:top]:
        print("%7d %5.1f%% %10d %6.1f  %s:%d  %s" % (s, 100.0 * s / max(total_s, 1), i, t / max(i, 1), p, ln, code[:110]))


if __name__ == "__main__":
    main()
