#!/usr/bin/env python
"""Per source-line-range (phase) breakdown of an `ncu --page source --csv --print-source cuda,sass` export of
k_graph_run: warp instructions, active lanes per instruction, stall samples.
usage: ncu_phase_breakdown.py src.csv graph_kernels.cuh"""
import collections
import csv
import re
import sys

csv.field_size_limit(10 ** 9)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    src = open(sys.argv[2]).read().splitlines()
    marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r"=+ (phase \d[^=(]*)", l)] if m]
    # the export is a sequence of blocks "File Path / Function Name / header / rows"; a row with a line number is a
    # source line (its SASS rows follow with an empty first column)
    want = sys.argv[2].split("/")[-1]
    fn = sys.argv[3] if len(sys.argv) > 3 else "k_graph_run<(bool)0>"
    per, other, h, on = collections.OrderedDict(), [0, 0, 0], None, False
    cur_file = cur_fn = ""
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1]
        elif r[0] == 'Function Name':
            cur_fn = r[1]
        elif r[0] == 'Line No':
            h = r
            il, ii, it, isamp = 0, h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
        elif h is not None and fn in cur_fn:
            try:
                L = int(r[il]); n = int(r[ii] or 0); t = int(r[it] or 0); sm = int(r[isamp] or 0)
            except ValueError:
                continue
            if cur_file.endswith(want):
                per[L] = (n, t, sm)
            else:
                other[0] += n; other[1] += t; other[2] += sm
    tw = sum(v[0] for v in per.values()) + other[0]; ts = sum(v[2] for v in per.values()) + other[2]
    print("other files (abs_rng.cuh draws, intrinsics headers): instr %.1f %%, lanes %.1f, samples %.1f %%"
          % (100.0 * other[0] / tw, other[1] / max(other[0], 1), 100.0 * other[2] / max(ts, 1)))
    W = sum(v[0] for v in per.values()); T = sum(v[1] for v in per.values()); S = sum(v[2] for v in per.values())
    print("total warp instr %.4g, active lanes %.1f, samples %d" % (W, T / max(W, 1), S))
    bounds = [(1, "helpers (inlined: draws, moves, pull_contact, load4 ...)")] + marks + [(len(src) + 1, "end")]
    for (a, name), (b, _) in zip(bounds[:-1], bounds[1:]):
        n = sum(v[0] for L, v in per.items() if a <= L < b); t = sum(v[1] for L, v in per.items() if a <= L < b)
        s = sum(v[2] for L, v in per.items() if a <= L < b)
        print("  lines %4d-%4d  %-60s instr %5.1f %%  lanes %4.1f  samples %5.1f %%" % (a, b - 1, name[:60], 100.0 * n / W, t / max(n, 1), 100.0 * s / max(S, 1)))
    print("  top lines by stall samples:")
    for L, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:22]:
        print("   L%-4d samples %5.1f %%  instr %5.1f %%  lanes %4.1f  %s" % (L, 100.0 * v[2] / S, 100.0 * v[0] / W, v[1] / max(v[0], 1), src[L - 1].strip()[:90] if L <= len(src) else ""))


if __name__ == "__main__":
    main()
