#!/bin/bash
# linearization variants (launch bounds / merged columns) built on the GPU box and timed one-stream (development aid)
run() {  # flags, threads, kpb
  DDPB_NVCC_EXTRA="$1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1 || { echo "build failed: $1"; return; }
  echo "== [$1] threads=$2 kpb=$3: $(DDPB_SUBBATCHES=1 DDPB_LIN_THREADS=$2 DDPB_LIN_KPB=$3 python tools/bwd_scan.py ${W:-c3} ${B:-1024} 2>&1 | grep linearize | sed 's/.*linearize/linearize/')"
}
run "" 256 3
run "-DDDPB_LIN_LB_THREADS=192 -DDDPB_LIN_LB_BLOCKS=2" 192 3
run "-DDDPB_LIN_LB_THREADS=192 -DDDPB_LIN_LB_BLOCKS=2" 192 2
run "-DDDPB_LIN_LB_THREADS=192 -DDDPB_LIN_LB_BLOCKS=2 -DDDPB_COL_MERGED=1" 192 3
run "-DDDPB_LIN_LB_THREADS=128 -DDDPB_LIN_LB_BLOCKS=2 -DDDPB_COL_MERGED=1" 128 3
run "-DDDPB_LIN_LB_THREADS=128 -DDDPB_LIN_LB_BLOCKS=3" 128 2
run "-DDDPB_LIN_LB_THREADS=128 -DDDPB_LIN_LB_BLOCKS=3 -DDDPB_COL_MERGED=1" 128 2
