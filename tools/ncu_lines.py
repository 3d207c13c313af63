#!/usr/bin/env python
"""Per-source-line totals of an ncu report with -lineinfo (instructions executed, stall samples).

  python tools/ncu_lines.py report.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines = []
    tot_i = tot_s = 0
    for r in rows:
        if len(r) > 8 and r[0].isdigit() and r[2] == "-":
            try:
                inst, samp = int(r[7] or 0), int(r[4] or 0)
            except ValueError:
                continue
            lines.append((inst, samp, int(r[0]), r[1].strip()))
            tot_i += inst; tot_s += samp
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    for inst, samp, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100.0 * inst / max(tot_i, 1):5.1f}% inst {100.0 * samp / max(tot_s, 1):5.1f}% samp  L{ln:<5} {src[:120]}")


if __name__ == "__main__":
    main()
