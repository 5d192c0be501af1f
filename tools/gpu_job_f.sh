#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "config4 or stream or 3d or final_cost or status" 2>&1 | tail -5 > gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
rm -f gpurun_out/r02f_config4_sweep.jsonl
for cpp in 1 2 4 8 16 32; do
  ICP_B200_STREAM_CELLS_PER_POINT=$cpp timeout 300 python tools/time_config4.py 296 8 >> gpurun_out/r02f_config4_sweep.jsonl 2>> gpurun_out/r02f_config4_sweep.err
done
cat gpurun_out/r02f_config4_sweep.jsonl | cut -c1-200
