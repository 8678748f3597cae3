#!/bin/bash
# launch-geometry sweep (development aid): threads per CTA of the linearization / backward kernels
for t in 64 96 128 192 256; do echo "LIN_THREADS=$t"; DDPB_LIN_THREADS=$t python tools/quick_time.py c3 | tail -1; done
for t in 32 64 128; do echo "c2 LIN_THREADS=$t"; DDPB_LIN_THREADS=$t python tools/quick_time.py c2 | tail -1; done
for t in 64 128 192 256; do echo "BWD_THREADS=$t"; DDPB_BWD_THREADS=$t python tools/quick_time.py c3 | tail -1; done
for t in 32 64; do echo "c2 BWD_THREADS=$t"; DDPB_BWD_THREADS=$t python tools/quick_time.py c2 | tail -1; done
