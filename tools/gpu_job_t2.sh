#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "overlapping_calls_other_kernels or back_to_back" 2>&1 | tail -12
