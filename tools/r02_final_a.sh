#!/bin/bash
# Final round-2 call, part A (run under gpurun, one GPU): GPU test suite, A/B of the value-based linearization (torque-affine
# points, fast tensor fill) against the previous library, per-phase cycles, ncu captures of the C5 linearization.
O=gpurun_out; V=gpurun_scratch/variants
timeout 300 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/f_pytest.log
for v in prev v3; do echo "[$v]"; DDPB_LIB=$V/libddpb_$v.so tools/bench_all.sh c5 c3; done
echo "[new]"; tools/bench_all.sh c5 c3
DDPB_LIB=$V/libddpb_pc.so timeout 120 python tools/phase_cycles.py c5 2>&1 | tail -1
DDPB_SUBBATCHES=1 timeout 300 ncu --set full --import-source on --clock-control none --kernel-name 'regex:^linearize_kernel$' --launch-skip 6 --launch-count 1 \
  -o $O/r02_full_c5_linearize_kernel -f python tools/bwd_scan.py c5 1024 > $O/r02_full_c5_linearize_kernel.log 2>&1; echo "ncu full rc $?"
M="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
KREG='^(linearize_kernel|linearize_pt_kernel|backward_kernel|backward_lg_kernel|ls_trials_kernel)'
DDPB_SUBBATCHES=1 timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:"$KREG" --csv --log-file $O/r02_metrics_c5.csv python tools/bwd_scan.py c5 1024 > /dev/null 2>&1; echo "ncu metrics rc $?"
