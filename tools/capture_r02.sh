#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU): ncu --set full of the main launch of every hot kernel for C2 and C3 with the
# solver on one stream, flop / DRAM-byte metrics of one solve's kernels, the launch list of the bench command.
# Outputs under gpurun_out/ (tools/summarize_profiles.py turns them into profiles/r02_*).
O=gpurun_out
KREG='^(linearize_kernel|linearize_pt_kernel|backward_kernel|backward_lg_kernel|ls_trials_kernel)'
for spec in c2:4096:linearize_pt_kernel:2 c2:4096:backward_lg_kernel:2 c2:4096:ls_trials_kernel:4 c3:1024:linearize_kernel:6 c3:1024:backward_kernel:2 c3:1024:ls_trials_kernel:4; do
  IFS=: read n b k skip <<< "$spec"
  DDPB_SUBBATCHES=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:^$k\$ --launch-skip $skip --launch-count 1 \
    -o $O/r02_full_${n}_$k -f python tools/bwd_scan.py $n $b > $O/r02_full_${n}_$k.log 2>&1
done
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
for w in c2:4096 c3:1024 c5:1024; do n=${w%:*}; b=${w#*:}
  DDPB_SUBBATCHES=1 timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:"$KREG" --csv \
    --log-file $O/r02_metrics_$n.csv python tools/bwd_scan.py $n $b > /dev/null 2>&1
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_b.log 2>&1 && \
DDPB_SUBBATCHES=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench_c2_onestream.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_ncu_bench1.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_bench_c2_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_ncu_bench4.log 2>&1
echo capture done
