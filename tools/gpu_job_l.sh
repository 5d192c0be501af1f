#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02l_pytest.log; tail -4 gpurun_out/r02l_pytest.log
timeout 300 python tools/time_target_update.py > gpurun_out/r02l_target_update.json 2>gpurun_out/r02l_target_update.err; cat gpurun_out/r02l_target_update.json; tail -2 gpurun_out/r02l_target_update.err
