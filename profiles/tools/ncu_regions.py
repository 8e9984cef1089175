"""Summarise an `ncu --set full --import-source on` capture of the tracking kernel by PHASE.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv > src.csv
    python profiles/tools/ncu_regions.py raw.csv src.csv <tracks in the launch>

The kernel's phases are separated by its CTA barriers, so the SASS between two BAR.SYNC instructions is one phase:
executed warp instructions per track and warp stall samples per phase, with the stall reasons of each."""
import collections
import csv
import sys

RAW = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
       "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
       "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    raw, src, ntr = sys.argv[1], sys.argv[2], int(sys.argv[3])
    names = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    r = list(csv.reader(open(raw)))
    h = r[0]
    print("== launch metrics ==")
    for i, n in enumerate(h):
        if n in RAW:
            print("%-68s %-16s %s" % (n, r[1][i], r[2][i]))
    r = list(csv.reader(open(src)))
    print("\n== %s ==" % r[0][1])
    h, rows = r[1], r[2:]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(x[isamp] or 0) for x in rows)
    totex = sum(int(x[iex] or 0) for x in rows)
    print("warp instructions executed %d (%.0f per track), warp stall samples %d" % (totex, totex / ntr, tot))
    regions, cur = [], dict(s=0, ex=0, start=0, st=collections.Counter())
    for k, x in enumerate(rows):
        cur["s"] += int(x[isamp] or 0)
        cur["ex"] += int(x[iex] or 0)
        for i in stall:
            if x[i] not in ("", "0"):
                cur["st"][h[i][6:]] += int(x[i])
        if "BAR.SYNC" in x[isrc]:
            cur["end"] = k
            regions.append(cur)
            cur = dict(s=0, ex=0, start=k + 1, st=collections.Counter())
    cur["end"] = len(rows) - 1
    regions.append(cur)
    print("\nphase (SASS rows)                         instr/track   share   samples  share   top stall reasons")
    for n, g in enumerate(regions):
        nm = names[n] if names and n < len(names) else "phase %d" % n
        t = sum(g["st"].values()) or 1
        top = ", ".join("%s %.0f%%" % (a, 100.0 * b / t) for a, b in g["st"].most_common(4))
        print("%-26s %5d-%5d   %10.0f  %5.1f%%   %7d  %5.1f%%   %s" % (nm, g["start"], g["end"], g["ex"] / ntr,
              100.0 * g["ex"] / totex, g["s"], 100.0 * g["s"] / tot, top))
    allst = collections.Counter()
    for g in regions:
        allst.update(g["st"])
    t = sum(allst.values())
    print("\nall phases: " + ", ".join("%s %.1f%%" % (a, 100.0 * b / t) for a, b in allst.most_common(10)))
    ops = collections.Counter()
    for x in rows:
        op = x[isrc].split()[0] if x[isrc].split() else ""
        if op.startswith("@"):
            op = x[isrc].split()[1]
        ops[op.split(".")[0]] += int(x[iex] or 0)
    print("instruction mix: " + ", ".join("%s %.1f%%" % (a, 100.0 * b / totex) for a, b in ops.most_common(14)))


if __name__ == "__main__":
    main()
