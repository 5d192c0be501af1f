#!/bin/bash
# round 2, job S: config 4 after the shrinking ball / own-cell-first ranking
mkdir -p gpurun_out
timeout 600 python tools/time_config4.py 1024 16 | cut -c1-600
timeout 900 python -m pytest tests -m gpu -q -x -k "config4 or 3d or fixed_point or streaming" 2>&1 | tail -3
