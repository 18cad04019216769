#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: per kernel, executed warp
instructions by source line and by opcode, plus stall samples.  usage: ncu_source_summary.py file.csv [kernel-substr] [agents]"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    agents = float(sys.argv[3]) if len(sys.argv) > 3 else 1e7
    rows = list(csv.reader(open(path)))
    kernels = collections.OrderedDict()
    cur = None
    hdr = None
    for r in rows:
        if r and r[0] == 'Function Name':
            cur = kernels.setdefault(r[1], [])
        elif r and r[0] == 'Line No':
            hdr = r
        elif cur is not None and len(r) > 10:
            cur.append(r)
    ci = hdr.index('Instructions Executed')
    cs = hdr.index('# Samples')
    stall_cols = [(k, n) for k, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
    for name, lines in kernels.items():
        if want not in name:
            continue
        sass = [r for r in lines if r[0] == '']
        tot = sum(I(r[ci]) for r in sass)
        samples = sum(I(r[cs]) for r in sass)
        print("== {}\n   warp instructions {}  = {:.1f} per 32 agents; samples {}".format(name[:90], tot, tot / (agents / 32), samples))
        by_line = collections.OrderedDict()
        line = None
        for r in lines:
            if r[0] != '':
                line = (r[0], r[1].strip()[:100])
                by_line.setdefault(line, [0, 0])
            else:
                by_line[line][0] += I(r[ci])
                by_line[line][1] += I(r[cs])
        top = sorted(by_line.items(), key=lambda kv: -kv[1][0])[:45]
        for (ln, src), (n, s) in top:
            print("   {:6.1f} instr/32ag {:5.1f}% samples  L{} {}".format(n / (agents / 32), 100.0 * s / max(1, samples), ln, src))
        ops = collections.Counter()
        for r in sass:
            t = r[3].split()
            op = t[1] if t and t[0].startswith('@') else (t[0] if t else '')
            ops[op.split('.')[0]] += I(r[ci])
        print("   opcodes:", ", ".join("{} {:.1f}".format(k, v / (agents / 32)) for k, v in ops.most_common(30)))
        st = collections.Counter()
        for r in sass:
            for k, n in stall_cols:
                st[n] += I(r[k])
        print("   stalls:", ", ".join("{} {:.1f}%".format(k, 100.0 * v / max(1, sum(st.values()))) for k, v in st.most_common(8)))


if __name__ == "__main__":
    main()
