#!/bin/bash
for cpp in 8; do ICP_B200_STREAM_CELLS_PER_POINT=$cpp timeout 300 python tools/time_config4.py 296 0 | cut -c1-140; done
echo "2 CTAs/SM"; for cpp in 8; do ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_m2 ICP_B200_STREAM_CELLS_PER_POINT=$cpp timeout 300 python tools/time_config4.py 592 8 | cut -c1-400; done
