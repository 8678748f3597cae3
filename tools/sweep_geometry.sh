#!/bin/bash
# launch-geometry sweep (development aid)
for k in 1 4 8 16; do for t in 32 64 128; do echo "c2 KPB=$k LIN_THREADS=$t"; DDPB_LIN_KPB=$k DDPB_LIN_THREADS=$t python tools/quick_time.py c2 | tail -1 | cut -c1-200; done; done
for k in 1 2; do for t in 64 128; do echo "c3 KPB=$k LIN_THREADS=$t"; DDPB_LIN_KPB=$k DDPB_LIN_THREADS=$t python tools/quick_time.py c3 | tail -1 | cut -c1-200; done; done
