#!/bin/bash
# bench lines of workloads for a library variant built with extra nvcc flags (development aid): var_bench.sh "<flags>" w1 w2 ...
flags="$1"; shift
DDPB_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > /dev/null 2>&1 || { echo "build failed"; exit 1; }
echo "== variant [$flags]"
tools/bench_all.sh "$@"
