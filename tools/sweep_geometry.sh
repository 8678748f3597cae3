#!/bin/bash
# launch-geometry sweep of the linearization kernel (development aid)
for k in 4 8 16; do for t in 64 128; do echo "c2 KPB=$k LIN_THREADS=$t"; DDPB_LIN_KPB=$k DDPB_LIN_THREADS=$t python tools/bwd_scan.py c2 4096 | tail -1 | sed 's/.*linearize/linearize/'; done; done
for k in 1 2 3; do for t in 64 128; do echo "c3 KPB=$k LIN_THREADS=$t"; DDPB_LIN_KPB=$k DDPB_LIN_THREADS=$t python tools/bwd_scan.py c3 1024 | tail -1 | sed 's/.*linearize/linearize/'; done; done
