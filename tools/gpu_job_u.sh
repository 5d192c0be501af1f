#!/bin/bash
# round 2, job U: config 4 after the fixed-point exit: grid sweep, ncu of the streaming kernel
mkdir -p gpurun_out
for c in 2 4 8 16; do ICP_B200_STREAM_CELLS_PER_POINT=$c timeout 300 python tools/time_config4.py 1024 0 | cut -c1-200; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_stream -s 1 -c 1 -o gpurun_out/r02u_stream python tools/time_config4.py 296 0 > gpurun_out/r02u_ncu_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep
