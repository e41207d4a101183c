#!/usr/bin/env python
"""Aggregate an ncu report's stall samples / executed warp instructions by source line and by opcode.
usage: python profiles/ncu_by_phase.py report.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def load(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", what, "--csv"],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    # ---- by CUDA source line (first kernel instance only) ----
    rows = load(rep, "cuda,sass")
    agg, cur, hdr, seen_kernel = collections.OrderedDict(), None, None, 0
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif r[0] != "" and hdr and r[0] not in ("Function Name",):
            try:
                key = (cur, int(r[0]))
                e = agg.setdefault(key, [0, 0, r[1].strip()])
                e[0] += int(r[isamp]); e[1] += int(r[iinst])
            except (ValueError, IndexError):
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print("total samples %d, warp instructions %d" % (tot, toti))
    print("samples%  instr%   file:line  source")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.1f%% %6.1f%%  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / toti, f, l, v[2][:100]))
    # ---- by opcode ----
    rows = load(rep, "sass")
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if hdr:
        ix = {h: i for i, h in enumerate(hdr)}
        op, ops = collections.Counter(), collections.Counter()
        for r in data:
            toks = r[ix["Source"]].split()
            o = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            op[o] += int(r[ix["Instructions Executed"]]); ops[o] += int(r[ix["# Samples"]])
        ti, ts = sum(op.values()) or 1, sum(ops.values()) or 1
        print("\nopcode      instr%%  samples%%   (%d SASS instructions in the kernel)" % len(data))
        for o, c in op.most_common(16):
            print("%-10s %6.1f%% %6.1f%%" % (o, 100 * c / ti, 100 * ops[o] / ts))


if __name__ == "__main__":
    main()
