#!/usr/bin/env python3
"""Developer tool: aggregate an ncu report's per-line warp-stall samples and shared-memory wavefronts of
t41rx_stream_rx_kernel by kernel section (function line ranges of rx_fast.cuh).
usage: tools/ncu_sections.py gpurun_out/prof.ncu-rep [n_stream_blocks]"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "t41_sdr_b200", "csrc", "rx_fast.cuh")


def sections():
    """(first_line, name) for every function-like block of rx_fast.cuh"""
    out = []
    pat = re.compile(r"^\s*(?:template <[^>]*>\s*)?__device__[^;]*?\b(\w+)\(")
    for n, line in enumerate(open(SRC), 1):
        m = pat.match(line)
        if m and "{" in line or (m and line.rstrip().endswith(",")):
            out.append((n, m.group(1)))
    return out


def main():
    rep = sys.argv[1]
    nsb = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    secs = sections()
    cur, ix = None, None
    agg = {}
    keys = ["# Samples", "stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_not_selected",
            "stall_selected", "stall_no_inst", "stall_mio", "stall_branch_resolving", "stall_dispatch", "stall_math",
            "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if r[0] == "Line No":
            ix = {h: i for i, h in enumerate(r)}
            continue
        if r[0].isdigit() and ix:
            ln = int(r[0])
            if cur == "rx_fast.cuh":
                name = "?"
                for first, nm in secs:
                    if first <= ln:
                        name = nm
            else:
                name = cur
            a = agg.setdefault(name, dict.fromkeys(keys, 0.0))
            for k in keys:
                try:
                    a[k] += float(r[ix[k]])
                except (ValueError, KeyError):
                    pass
    tot = sum(a["# Samples"] for a in agg.values())
    print("%-22s %7s %6s | %6s %6s %6s %6s %6s %6s %6s %6s | %9s %8s %8s" % (
        "section", "samples", "%", "barr", "ssb", "lsb", "wait", "nsel", "sel", "noinst", "mio", "instr/sb", "smemwf/sb", "excess/sb"))
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["# Samples"]):
        if a["# Samples"] < 0.002 * tot and a["Instructions Executed"] < 1:
            continue
        print("%-22s %7.0f %5.1f%% | %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f | %9.1f %8.1f %8.1f" % (
            name[:22], a["# Samples"], 100 * a["# Samples"] / tot, a["stall_barrier"], a["stall_short_sb"], a["stall_long_sb"],
            a["stall_wait"], a["stall_not_selected"], a["stall_selected"], a["stall_no_inst"], a["stall_mio"],
            a["Instructions Executed"] / nsb, a["L1 Wavefronts Shared"] / nsb, a["L1 Wavefronts Shared Excessive"] / nsb))
    print("total samples %.0f; instr/stream-block %.0f; smem wavefronts/stream-block %.0f (excess %.0f)" % (
        tot, sum(a["Instructions Executed"] for a in agg.values()) / nsb,
        sum(a["L1 Wavefronts Shared"] for a in agg.values()) / nsb,
        sum(a["L1 Wavefronts Shared Excessive"] for a in agg.values()) / nsb))


if __name__ == "__main__":
    main()
