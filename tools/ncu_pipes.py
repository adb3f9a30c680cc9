"""Pipe / issue / stall summary of one kernel from an `ncu --set full` capture (raw page as CSV):
    ncu -i capture.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_pipes.py raw.csv > profiles/<name>.txt"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("kernel:", r[ix["Kernel Name"]][:100], " grid", r[ix["Grid Size"]], " block", r[ix["Block Size"]])
    for k in want:
        if k in ix:
            print(f"  {k:95s} {r[ix[k]]:>16s} {units[ix[k]]}")
    st = sorted(((float(r[i].replace(',', '')), h) for h, i in ix.items()
                 if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")), reverse=True)
    print("  warp stall reasons (warps per issue-active cycle):")
    for v, h in st[:8]:
        print(f"    {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:6.2f}")
