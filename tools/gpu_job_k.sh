#!/bin/bash
for rep in 1 2; do
for v in lib lib_ex; do
  ICP_B200_LIB_DIR=$PWD/icp_rust_b200/$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['e2e']['value'], d['clocks'])"
done; done
