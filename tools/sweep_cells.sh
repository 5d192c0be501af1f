#!/bin/bash
# cell-budget sweeps: streaming kernel's per-pair grids (config-4 shape) and the lean kernel's shared-memory grid (config 2)
for c in ${STREAM_CPP:-2 8 16 32 64}; do
  echo "stream cells per point $c: $(ICP_B200_STREAM_CELLS_PER_POINT=$c python tools/time_config4.py 296 2>&1 | tail -2 | head -1)"
done
for b in ${LEAN_BUDGET:-2048 4096 8192 12288}; do
  echo "lean cell budget $b: $(ICP_B200_CELL_BUDGET=$b python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['config']['kernel'])")"
done
