#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples and executed instructions by CUDA source line.
usage: python profiles/ncu_by_line.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    data, cur, hdr = [], None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif r[0] != "" and hdr and r[0] != "Function Name":
            try:
                data.append((int(r[isamp]), int(r[iinst]), cur, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    toti = sum(d[1] for d in data) or 1
    print("total samples %d, warp instructions %d" % (tot, toti))
    print("samples%  instr%   file:line  source")
    for d in sorted(data, reverse=True)[:top]:
        print("%6.1f%% %6.1f%%  %s:%d  %s" % (100 * d[0] / tot, 100 * d[1] / toti, d[2], d[3], d[4][:110]))


if __name__ == "__main__":
    main()
