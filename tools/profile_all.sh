#!/bin/bash
# Round-2 ncu evidence, run on the GPU box under gpurun (one call): for each dominant kernel a plain run of
# tools/profile_cmd.py, then the same command under `ncu --set full`; the reports (60-75 MB each: they embed the whole
# cubin) stay in /tmp, what travels back are the text summaries and the facts JSON bench.py reads.
#   gpurun -- 'bash tools/profile_all.sh'
set -u
OUT=gpurun_out/r02_prof
mkdir -p $OUT
python tools/profile_cmd.py c3 c2 c4d c5 > $OUT/plain_run.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_run.log; exit 1; }
# which launch of the command is the measured one of each kernel (see tools/profile_cmd.py: warm-up + measured launch each)
run() {  # name  workload  kernel-regex  launches-to-skip
  ncu --set full --import-source on --clock-control none -k regex:"$3" -c 1 -s $4 -o /tmp/prof_$1 -f python tools/profile_cmd.py $2 > $OUT/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep 30 > $OUT/r02_prof_$1_summary.txt 2>&1
}
# c3: argmax calls of this size run in two passes (pass 1 = hi*hi MMAs only over everything, marking kernel, pass 2 = the full
# form over the listed items); per basis the command makes a warm-up and a measured call, so the measured pass 1 is rff_fused
# launch 2 (fp32 form) / 6 (fast form).  The full-form kernel over everything (= what pass 2 runs, and the single pass
# TKBO_PRUNE=0 selects) is captured as before.
run k1_c3_fp32_pass1 c3 rff_fused 2
run k1_c3_fast_pass1 c3 rff_fused 6
export TKBO_PRUNE=0
run k1_c3_fp32 c3 rff_fused 1
run k1_c3_fast c3 rff_fused 3
unset TKBO_PRUNE
run k1_c2 c2 rff_fused 1
run k1_c4d c4d rff_fused 1
run k2_c5 c5 mpes_all_perms 1
python tools/ncu_facts.py $OUT/r02_ncu_facts.json /tmp/prof_k1_c3_fp32_pass1.ncu-rep /tmp/prof_k1_c3_fast_pass1.ncu-rep /tmp/prof_k1_c3_fp32.ncu-rep /tmp/prof_k1_c3_fast.ncu-rep /tmp/prof_k1_c2.ncu-rep /tmp/prof_k1_c4d.ncu-rep /tmp/prof_k2_c5.ncu-rep > $OUT/ncu_facts.log 2>&1
cat $OUT/plain_run.log
grep -h "time_duration\|tensor_cycles_active\|dram__bytes_read" $OUT/r02_prof_*_summary.txt
