import csv,sys,subprocess,collections,re
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,r=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__cycles_elapsed.avg','sm__issue_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_srcunit_tex.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__cycles_active.avg','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active']
for i,h in enumerate(hdr):
    if h in want: print(f"{h} = {r[i]} {units[i]}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=[i for i,x in enumerate(rows) if x and x[0]=="Address"]
h=rows[hi[0]]; body=rows[hi[0]+1:(hi[1]-1 if len(hi)>1 else len(rows))]
ci={n:i for i,n in enumerate(h)}
base=int(body[0][ci["Address"]],16)
tot=sum(int(x[ci["# Samples"]] or 0) for x in body)
print("total samples",tot)
for x in sorted(body,key=lambda x:-int(x[ci["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    st={k:int(x[ci[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}
    main=sorted(st.items(), key=lambda kv:-kv[1])[:3]
    print(hex(int(x[ci["Address"]],16)-base), "%6d %5.1f%%"%(int(x[ci["# Samples"]]),100*int(x[ci["# Samples"]])/tot), x[ci["Source"]][:64].ljust(64), main)
