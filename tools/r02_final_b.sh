#!/bin/bash
# Final round-2 call, part B (run under gpurun, one GPU): GPU test suite on the product library, value-mode tests and a C5 bench
# line for the looped-fill variant (-DDDPB_FILL_LOOP=1), bench lines of C2 (default), C3, C4, C5, ncu full capture of the
# C5 linearization of whichever library is faster.
O=gpurun_out; V=gpurun_scratch/variants
timeout 300 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc $?"; tail -1 $O/f_pytest.log
DDPB_LIB=$V/libddpb_loop.so timeout 200 python -m pytest tests -m gpu -x -q -k "fdvalues or c5" > $O/f_pytest_loop.log 2>&1; echo "pytest[loop] rc $?"; tail -1 $O/f_pytest_loop.log
echo "[u7]"; DDPB_LIB=$V/libddpb_u7.so tools/bench_all.sh c5
timeout 200 python bench.py --workload c5 --steps 3 --warmup 3 > $O/r02_bench_c5_v3.json 2> $O/r02_bench_c5_v3.err
DDPB_LIB=$V/libddpb_loop.so timeout 200 python bench.py --workload c5 --steps 3 --warmup 3 > $O/r02_bench_c5_loop.json 2> $O/r02_bench_c5_loop.err
for v in v3 loop; do echo "[$v]"; python tools/bench_brief.py < $O/r02_bench_c5_$v.json; done
W=$(python - <<'PY'
import json
v=lambda n: json.loads(open(f"gpurun_out/r02_bench_c5_{n}.json").read().strip().splitlines()[-1])["value"]
try: print("loop" if v("loop") > v("v3") else "v3")
except Exception: print("v3")
PY
)
echo "winner $W" | tee $O/winner.txt
L=ddp-pinocchio_b200/libddpb.so; [ "$W" = loop ] && L=$V/libddpb_loop.so
DDPB_LIB=$L DDPB_SUBBATCHES=1 timeout 200 ncu --set full --import-source on --clock-control none --kernel-name 'regex:^linearize_kernel$' --launch-skip 6 --launch-count 1 \
  -o $O/r02_full_c5_linearize_kernel -f python tools/bwd_scan.py c5 1024 > $O/r02_full_c5_linearize_kernel.log 2>&1; echo "ncu full rc $?"
timeout 200 python bench.py > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err; python tools/bench_brief.py < $O/r02_bench_c2.json
for w in c3 c4; do timeout 200 python bench.py --workload $w --steps 3 --warmup 3 > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; python tools/bench_brief.py < $O/r02_bench_$w.json; done
