#!/bin/bash
# Linearization launch-geometry sweep (development aid): threads per CTA x knots per CTA, one-stream timings of tools/bwd_scan.py
w=${1:-c3}; B=${2:-1024}
GEOMS=${GEOMS:-256:2 160:2 128:2 192:2 256:3 224:3 128:1 96:1}
for cfg in $GEOMS; do
  t=${cfg%%:*}; k=${cfg##*:}
  echo "== threads=$t kpb=$k: $(DDPB_SUBBATCHES=1 DDPB_LIN_THREADS=$t DDPB_LIN_KPB=$k python tools/bwd_scan.py $w $B 2>&1 | grep linearize | sed 's/.*linearize/linearize/')"
done
