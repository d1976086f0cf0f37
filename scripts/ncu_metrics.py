import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_tensor','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','launch__occupancy_limit','sm__maximum_warps_per_active_cycle_pct','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum']
for r in rows[2:]:
    print("----")
    for i,h in enumerate(hdr):
        if any(h==w or h.startswith(w) for w in want):
            print(f"  {h} [{rows[1][i]}] = {r[i]}")
