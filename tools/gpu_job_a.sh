#!/bin/bash
# round 2, GPU job A: parity suite, bench line, census, phase clocks of the lean kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 900 python tools/census.py 131072 > gpurun_out/r02a_census.json 2> gpurun_out/r02a_census.err
ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_timing timeout 300 python tools/profile_fused.py 592 2 > gpurun_out/r02a_lean_timing.log 2>&1
tail -5 gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_census.json
tail -3 gpurun_out/r02a_lean_timing.log
