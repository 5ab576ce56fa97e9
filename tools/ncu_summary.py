#!/usr/bin/env python3
"""Text summary of an .ncu-rep (ncu --set full): one block per kernel launch with the counters
the roofline discussion needs, plus the top stall reasons.  usage: ncu_summary.py <rep>"""
import csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (fp32/int) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread instr"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread instr"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread instr"),
]
for r in body:
    print("=" * 100)
    print(r[ix["Kernel Name"]][:110])
    for k, label in KEYS:
        if k in ix:
            print("  %-46s %s %s" % (label, r[ix[k]], units[ix[k]]))
    st = []
    for h, i in ix.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  warps stalled per issue (top): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
