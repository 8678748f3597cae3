#!/bin/bash
# Refresh of the linearization evidence after a linearization-only change (run under gpurun): ncu --set full of its main
# launch for C2 and C3, the per-kernel metric pass, bench lines and the one-stream launch list of the bench command.
set -x
O=gpurun_out
for w in c2:4096 c3:1024; do n=${w%:*}; b=${w#*:}
  DDPB_SUBBATCHES=1 timeout 200 ncu --set full --import-source on --clock-control none --kernel-name regex:^linearize_kernel --launch-skip 12 --launch-count 1 \
    -o $O/full_${n}_linearize_kernel -f python tools/bwd_scan.py $n $b > $O/full_${n}_linearize_kernel.log 2>&1
done
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
for w in c2:4096 c3:1024 c5:1024; do n=${w%:*}; b=${w#*:}
  DDPB_SUBBATCHES=1 timeout 200 ncu --metrics $M --clock-control none --kernel-name regex:'^(linearize_kernel|backward_kernel|forward_kernel|cost_old_kernel)' --launch-skip 12 --launch-count 6 --csv \
    --log-file $O/metrics_$n.csv python tools/bwd_scan.py $n $b > /dev/null 2>&1
done
timeout 300 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --workload c3 --steps 10 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --workload c5 --steps 3 > $O/bench_c5.json 2> $O/bench_c5.err
DDPB_SUBBATCHES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bench_c2_onestream.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench1.log 2>&1
timeout 300 python bench.py --workload c4 --steps 5 > $O/bench_c4.json 2> $O/bench_c4.err
echo capture done
