#!/bin/bash
# times the device path of bench.py on each kernel instance able to hold a 1080-point scan
for i in ${SWEEP_IDS:-3 4 5}; do
  echo "== instance $i"; ICP_B200_FORCE_INSTANCE=$i timeout 300 python bench.py --no-cpu-baseline --steps 6 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['kernel'], d['e2e']['value'], d.get('ordered_sums_mode'))"
done
