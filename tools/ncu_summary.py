#!/usr/bin/env python
"""Compact summary of an .ncu-rep (read here, on the CPU box):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep

Prints one block per captured launch with the metrics the roofline discussion needs: duration, DRAM traffic,
occupancy, registers, shared memory, pipe utilisation (fp64 / alu / fma / lsu / xu), issue slot use and the top
warp-stall reasons.  `--source N` adds the N hottest source lines (needs -lineinfo + --import-source on).
"""

import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_warps", "occ limit warps (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 % (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "pipe fp64 cycles %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "pipe dmma %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-inst"),
    ("sass__inst_executed_local_loads", "local loads"),
    ("sass__inst_executed_local_stores", "local stores"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    nsrc = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
    hdr, units, launches = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    for row in launches:
        name = row[idx["Kernel Name"]]
        print(f"### {name[:110]}")
        for k, label in KEYS:
            if k in idx:
                print(f"- {label}: {row[idx[k]]} {units[idx[k]]}")
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(row[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_") : -len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("- top stalls (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
        print()
    if nsrc:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if rows:
            h = rows[0]
            try:
                si = h.index("Source")
                wi = [i for i, x in enumerate(h) if x.startswith("Warp Stall Sampling (All")][0]
                ii = [i for i, x in enumerate(h) if x.startswith("Instructions Executed")][0]
            except (ValueError, IndexError):
                print("source page: columns not found", h[:12])
                return
            body = []
            for r in rows[1:]:
                try:
                    body.append((float(r[wi].replace(",", "") or 0), float(r[ii].replace(",", "") or 0), r[si]))
                except (ValueError, IndexError):
                    pass
            tot = sum(b[0] for b in body) or 1.0
            body.sort(reverse=True)
            print(f"### hottest source lines (stall samples, share of {tot:.0f})")
            for s, n, src in body[:nsrc]:
                print(f"- {s / tot:6.3f}  inst={n:12.0f}  {src.strip()[:130]}")


if __name__ == "__main__":
    main()
