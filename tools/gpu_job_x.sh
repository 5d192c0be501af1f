#!/bin/bash
# round 2, job X: config 4 with the binary32 record filter
mkdir -p gpurun_out
timeout 300 python tools/time_config4.py 1024 16 | cut -c1-500
ICP_B200_STREAM_NN_F32=0 timeout 300 python tools/time_config4.py 1024 0 | cut -c1-200
for c in 2 8; do ICP_B200_STREAM_CELLS_PER_POINT=$c timeout 300 python tools/time_config4.py 1024 0 | cut -c1-200; done
timeout 900 python -m pytest tests -m gpu -q -x -k "config4 or 3d or streaming or fixed_point or l_shape" 2>&1 | tail -2
