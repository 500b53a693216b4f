#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/:
key metrics per kernel, the per-opcode executed-instruction mix and the most-stalled SASS lines.
usage: summarize_ncu.py report.ncu-rep [warp_blocks_per_launch] > summary.txt   (runs on the CPU box)"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    warp_blocks = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        print("  stall reasons (warps per issue-active cycle):")
        for i, k in enumerate(hdr):
            if "issue_stalled" in k and "per_issue_active" in k and float(r[i]) > 0.05:
                print("     %-28s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i]))
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    # the source page holds one table per kernel: "Kernel Name" line, header line, instructions
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            name, hdr = src[i][1], src[i + 1]
            j = i + 2
            body = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                if len(src[j]) == len(hdr):
                    body.append(src[j])
                j += 1
            iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            ops, samp, tot, tots = collections.Counter(), collections.Counter(), 0, 0
            for r in body:
                t = r[iS].split()
                op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                ops[op] += int(r[iE])
                samp[op] += int(r[iN])
                tot += int(r[iE])
                tots += int(r[iN])
            print("=" * 100)
            print("SASS mix:", name, "-- warp instructions executed:", tot)
            if warp_blocks:
                print("  per warp-block (32 coefficient blocks): %.1f" % (tot / warp_blocks))
            for op, c in ops.most_common(24):
                extra = "  %8.1f per warp-block" % (c / warp_blocks) if warp_blocks else ""
                print("  %-10s %14d  %5.1f%% of instructions  %5.1f%% of stall samples%s" % (op, c, 100.0 * c / tot, 100.0 * samp[op] / max(1, tots), extra))
            stall = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            print("  most-sampled instructions:")
            for r in sorted(body, key=lambda r: -int(r[iN]))[:16]:
                top = sorted(((int(r[k]), h) for k, h in stall if r[k] not in ("", "0")), reverse=True)[:2]
                print("   %5.2f%%  %-72s %s" % (100.0 * int(r[iN]) / max(1, tots), r[iS].strip()[:72], ", ".join(f"{h}={v}" for v, h in top)))
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main()
