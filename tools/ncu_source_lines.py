#!/usr/bin/env python3
"""Aggregate an .ncu-rep source page (cuda,sass view) per source line: instructions executed + stall samples.

    python tools/ncu_source_lines.py report.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath = None
    per = {}
    tot_inst = tot_samp = 0
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None or r[0] == "":
            continue  # SASS rows are summed into their source line row already
        try:
            inst = int(r[i_inst])
            samp = int(r[i_samp])
        except (ValueError, IndexError):
            continue
        key = (fpath, int(r[0]), r[1].strip()[:90])
        a = per.setdefault(key, [0, 0])
        a[0] += inst
        a[1] += samp
        tot_inst += inst
        tot_samp += samp
    print(f"total warp-instructions {tot_inst:,}  samples {tot_samp:,}")
    print("-- by instructions executed")
    for key, (inst, samp) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * inst / max(1, tot_inst):5.1f}% inst {100 * samp / max(1, tot_samp):5.1f}% samp  {key[0]}:{key[1]}  {key[2]}")
    print("-- by stall samples")
    for key, (inst, samp) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * samp / max(1, tot_samp):5.1f}% samp {100 * inst / max(1, tot_inst):5.1f}% inst  {key[0]}:{key[1]}  {key[2]}")


if __name__ == "__main__":
    main()
