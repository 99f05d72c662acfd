#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into the text summary committed here.

    python profiles/summarize.py gpurun_out/prof_x.ncu-rep "title" > profiles/r01_x.md

Reads the report with `ncu -i ... --page raw/source --csv`; nothing here runs on a GPU."""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    print(f"# {title}\n\nsource report: `{rep}` (scratch, not committed); numbers under the profiler are "
          "for attribution only, never bench values.\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}\n\n| metric | value | unit |\n|---|---|---|")
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        for k in WANT:
            if k in d and d[k] != "":
                print(f"| {k} | {d[k]} | {u[k]} |")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                         and v not in ("", "n/a")), reverse=True)[:8]
        if stalls:
            print("\nwarp stall reasons (warps per issue-active cycle): " +
                  ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls))
        print()
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv"))))
    if len(src) > 2 and "Instructions Executed" in src[1]:
        h = src[1]
        iS, iE, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        ops, thr = collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) <= iT:
                continue
            if not (r[iE] or "0").isdigit():
                continue   # (reports with several kernels repeat the header rows)
            s = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
            op = ".".join(s.split()[0].split(".")[:3]) if s else "?"
            ops[op] += int(r[iE] or 0)
            thr[op] += int(r[iT] or 0)
        tot = sum(ops.values())
        print(f"## SASS opcode mix (warp instructions executed, total {tot / 1e9:.2f} G)\n\n"
              "| opcode | G inst | % | threads/inst |\n|---|---|---|---|")
        for op, c in ops.most_common(24):
            print(f"| {op} | {c / 1e9:.2f} | {100 * c / tot:.1f} | {thr[op] / max(c, 1):.1f} |")
        tags = [op for op in ops if op.startswith(("UBLKCP", "UTMA", "SYNCS", "MUFU", "DFMA", "TCGEN", "UTC"))]
        print("\nSASS mnemonics present (TMA bulk copy = UBLKCP, mbarrier = SYNCS): " + ", ".join(sorted(tags)))
    cs = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, out, hdr2 = None, [], None
    for r in cs:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r and r[0].isdigit() and "Instructions Executed" in hdr2:
            try:
                out.append((cur, int(r[0]), r[1], int(r[hdr2.index("Instructions Executed")] or 0),
                            int(r[hdr2.index("# Samples")] or 0)))
            except ValueError:
                pass
    if out:
        tot = sum(o[3] for o in out) or 1
        ts = sum(o[4] for o in out) or 1
        print("\n## hottest source lines (-lineinfo)\n\n| file:line | G inst | % inst | % stall samples | source |\n|---|---|---|---|---|")
        for f, l, s, e, sm in sorted(out, key=lambda o: -o[3])[:22]:
            print(f"| {f}:{l} | {e / 1e9:.2f} | {100 * e / tot:.1f} | {100 * sm / ts:.1f} | `{s.strip()[:90]}` |")


if __name__ == "__main__":
    main()
