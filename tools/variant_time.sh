#!/bin/bash
# build a variant of the library with extra nvcc flags on the GPU box and time the clean passes (development aid)
# usage: tools/variant_time.sh "<nvcc flags>" workload...
flags="$1"; shift
DDPB_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" || exit 1
echo "== variant [$flags]"
python tools/bwd_diag.py "$@" 2>&1 | grep -v "line search"
