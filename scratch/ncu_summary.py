"""Summarise an .ncu-rep: key raw metrics + stall samples grouped by warp role (SASS landmarks)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
print("metric,unit,value")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%s,%s,%s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
tot = sum(int(r[si] or 0) for r in data)
print("# total samples", tot, "sass instructions", len(data))
w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
for a in range(0, len(data), w):
    blk = data[a:a + w]
    s = sum(int(r[si] or 0) for r in blk)
    ex = sum(int(r[ie] or 0) for r in blk)
    marks = set()
    for r in blk:
        for k in ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "USETMAXREG", "SYNCS", "UTCBAR", "NANOSLEEP", "EXIT", "STS", "STG", "MUFU", "F2FP", "FFMA", "FADD"]:
            if k in r[1]:
                marks.add(k)
    print("# sass[%5d:%5d] samples %6d (%4.1f%%) executed %10d %s" % (a, a + w, s, 100.0 * s / max(tot, 1), ex, " ".join(sorted(marks))))
