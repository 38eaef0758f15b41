#!/usr/bin/env python3
"""Turns the scratch artefacts of a GPU run (gpurun_out/) into the committed summaries under profiles/.

    python tools/make_profiles.py TAG [--rep gpurun_out/prof.ncu-rep --kernel-regex k_assemble_ws]
                                      [--launches gpurun_out/launches.csv] [--sass KERNEL_MANGLED_SUBSTRING]

Runs in the authoring container (needs ncu / cuobjdump, no GPU).  Developer tool."""
import argparse
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def raw_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    text, traffic = [], {}
    for vals in rows[2:]:
        rec = dict(zip(hdr, zip(units, vals)))
        name = rec["Kernel Name"][1]
        text.append("---- " + name[:110])
        for k in KEYS:
            if k in rec:
                text.append("  %-90s %s %s" % (k, rec[k][1], rec[k][0]))
        for k in hdr:
            m = STALLS.match(k)
            if m and float(rec[k][1] or 0) >= 0.05:
                text.append("  %-90s %s" % (k, rec[k][1]))
        def gb(k):
            u, v = rec[k]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
        traffic[name] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    return "\n".join(text) + "\n", traffic


def launches_summary(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(rows[start])}
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        if r[ix["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        elif r[ix["Metric Unit"]] in ("ms", "msecond"):
            ns *= 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    lines = ["%-82s %5s %10s %9s %6s" % ("kernel", "n", "total ms", "avg ms", "share")]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-82s %5d %10.3f %9.4f %5.1f%%" % (k[:82], n, t / 1e6, t / n / 1e6, 100 * t / tot))
    ours = sum(v[0] for k, v in agg.items() if "fe::" in k)
    lines.append("launches: %d, of which fe:: kernels %d; total device time %.3f ms" % (sum(v[0] for v in agg.values()), ours, tot / 1e6))
    return "\n".join(lines) + "\n"


def sass_excerpt(lib, needle, patterns=("UBLKCP", "SYNCS", "LDGSTS", "USETMAXREG", "UTMA", "FENCE", "DEPBAR", "LDGDEPBAR")):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines, keep, name = [], False, None
    counts = collections.Counter()
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            keep = needle in m.group(1)
            name = m.group(1)
            if keep:
                lines.append("Function : " + name)
            continue
        if not keep:
            continue
        ins = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
        if not ins:
            continue
        op = ins.group(2).split()[1] if ins.group(2).startswith("@") else ins.group(2).split()[0]
        counts[op.split(".")[0]] += 1
        if any(p in ins.group(2) for p in patterns):
            lines.append("  /*%s*/  %s" % (ins.group(1), ins.group(2)))
    lines.append("opcode histogram: " + ", ".join("%s %d" % kv for kv in counts.most_common(40)))
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--launches")
    ap.add_argument("--sass")
    ap.add_argument("--name", default="")
    a = ap.parse_args()
    prof = os.path.join(ROOT, "profiles")
    if a.rep:
        text, traffic = "", {}
        for rep in a.rep:
            t, tr = raw_summary(rep)
            text += t
            traffic.update(tr)
        open(os.path.join(prof, "%s_ncu_full_%s.txt" % (a.tag, a.name or "kernels")), "w").write(text)
        print(json.dumps(traffic, indent=1))
    if a.launches:
        open(os.path.join(prof, "%s_launches_%s.txt" % (a.tag, a.name or "bench")), "w").write(launches_summary(a.launches))
    if a.sass:
        open(os.path.join(prof, "%s_sass_%s.txt" % (a.tag, a.name or "kernel")), "w").write(
            sass_excerpt(os.path.join(ROOT, "feprogram_b200", "libfeprogram_b200.so"), a.sass))


if __name__ == "__main__":
    main()
