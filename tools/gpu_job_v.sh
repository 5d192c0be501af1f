#!/bin/bash
# round 2, job V: ncu of the streaming kernel on config 4 (main launch of the second call), timing at the new default grid
mkdir -p gpurun_out
timeout 300 python tools/time_config4.py 1024 16 | cut -c1-500
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_stream -s 2 -c 1 -o gpurun_out/r02v_stream python tools/time_config4.py 296 0 > gpurun_out/r02v_ncu_stream.log 2>&1
tail -2 gpurun_out/r02v_ncu_stream.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -q -x -k "config4 or 3d or streaming or fixed_point" 2>&1 | tail -2
