import csv, io, subprocess, sys
path=sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = {"Kernel Name":"name","Grid Size":"grid","Block Size":"blk","gpu__time_duration.sum":"us","dram__bytes_read.sum":"rdB","dram__bytes_write.sum":"wrB",
 "dram__throughput.avg.pct_of_peak_sustained_elapsed":"dram%","sm__throughput.avg.pct_of_peak_sustained_elapsed":"sm%",
 "l1tex__throughput.avg.pct_of_peak_sustained_elapsed":"l1%","lts__throughput.avg.pct_of_peak_sustained_elapsed":"l2%",
 "sm__warps_active.avg.pct_of_peak_sustained_active":"occ%","launch__registers_per_thread":"regs",
 "smsp__issue_active.avg.pct_of_peak_sustained_active":"issue%","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active":"lsu%",
 "smsp__inst_executed.sum":"inst","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed":"shwave%",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio":"st_long","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio":"st_short",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio":"st_wait","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio":"st_math",
 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio":"st_bar","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio":"st_lg",
 "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio":"st_mio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio":"st_nsel",
 "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed":"tensor%"}
idx = {k: hdr.index(k) for k in want if k in hdr}
for r in rows[2:]:
    d = {want[k]: r[i] for k,i in idx.items()}
    name = d.pop("name").replace("void ","").replace("sg::","").replace("<unnamed>::","")[:48]
    print(name, " ".join(f"{k}={v[:9]}" for k,v in d.items()))
