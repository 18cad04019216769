#!/usr/bin/env python
"""Condenses ncu exports into the tables committed under profiles/.

    python scripts/make_profile_summary.py <tag> <raw.csv> <sass-source.csv> [agents]

raw.csv:  ncu -i X.ncu-rep --page raw --csv
sass.csv: ncu -i X.ncu-rep --page source --csv --print-source sass
Writes profiles/<tag>_ncu_full_raw.csv (the selected metrics), profiles/<tag>_instruction_mix.md (per kernel: executed
warp instructions per 32 agents by opcode, stall reasons, the 12 instructions with the most stall samples).
"""
import collections
import csv
import os
import sys

csv.field_size_limit(10 ** 9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct")


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    tag, raw, sass = sys.argv[1], sys.argv[2], sys.argv[3]
    agents = float(sys.argv[4]) if len(sys.argv) > 4 else 1e7
    rows = list(csv.reader(open(raw)))
    h = rows[0]
    idx = [h.index(k) for k in KEEP if k in h]
    with open(os.path.join(ROOT, "profiles", tag + "_ncu_full_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    launches_of = collections.Counter(r[h.index("Kernel Name")] for r in rows[2:])
    inst_of = {}    # executed warp instructions of ONE launch, from the raw page (the source page sums over the launches)
    if "smsp__inst_executed.sum" in h:
        for r in rows[2:]:
            inst_of.setdefault(r[h.index("Kernel Name")], float(r[h.index("smsp__inst_executed.sum")] or 0))
    srows = list(csv.reader(open(sass)))
    kernels, cur, hdr = collections.OrderedDict(), None, None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = kernels.setdefault(r[1], [])
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and len(r) > 10:
            cur.append(r)
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    out = ["# {}: instruction mix and stall samples from the ncu source page (SASS)\n".format(tag),
           "Counts are executed WARP instructions per launch divided by agents / 32 ({:.0f} agents): the instructions a warp "
           "issues per 32 agents, spin-wait iterations included.\n".format(agents)]
    for name, lines in kernels.items():
        short = name.split("(")[0].replace("cabs::", "").replace("void ", "")
        base = short.split("<")[0]
        nl = max(1, sum(v for k, v in launches_of.items() if base in k))
        one = [v for k, v in inst_of.items() if base in k and v > 0]
        if one:
            nl = max(1, round(sum(I(r[ci]) for r in lines) / one[0]))
        tot = sum(I(r[ci]) for r in lines) / nl
        samples = sum(I(r[cs]) for r in lines)
        ops = collections.Counter()
        for r in lines:
            t = r[1].split()
            op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
            ops[op.split(".")[0]] += I(r[ci]) / nl
        st = collections.Counter()
        for r in lines:
            for i, n in stall:
                st[n] += I(r[i])
        out.append("\n## {}\n".format(short))
        out.append("{:.0f} warp instructions per launch = **{:.0f} per 32 agents**; {} stall samples\n".format(
            tot, tot / (agents / 32), samples))
        out.append("opcodes (per 32 agents): " + ", ".join("{} {:.1f}".format(k, v / (agents / 32)) for k, v in ops.most_common(18)) + "\n")
        ssum = max(1, sum(st.values()))
        out.append("stall reasons: " + ", ".join("{} {:.1f} %".format(k.replace("stall_", ""), 100.0 * v / ssum) for k, v in st.most_common(7)) + "\n")
        out.append("\n| stall samples | executed | instruction | dominant reason |\n|---|---|---|---|")
        seen = set()
        for r in sorted(lines, key=lambda r: -I(r[cs])):
            key = (r[0], r[1])
            if key in seen:
                continue
            seen.add(key)
            if len(seen) > 12:
                break
            why = max(((I(r[i]), n) for i, n in stall))[1].replace("stall_", "")
            out.append("| {} | {} | `{}` | {} |".format(I(r[cs]), I(r[ci]), r[1].strip()[:70], why))
        out.append("")
    with open(os.path.join(ROOT, "profiles", tag + "_instruction_mix.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote profiles/{0}_ncu_full_raw.csv, profiles/{0}_instruction_mix.md".format(tag))


if __name__ == "__main__":
    main()
