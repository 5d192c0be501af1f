#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_stream -s 2 -c 1 -o gpurun_out/r02i_stream python tools/time_config4.py 148 0 > gpurun_out/r02i_ncu_stream.log 2>&1
ls -la gpurun_out/r02i_stream.ncu-rep
