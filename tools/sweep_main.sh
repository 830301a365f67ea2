#!/bin/bash
# main-kernel tile sweep on C4 (tuning aid)
for cfg in "16 2" "8 3" "8 4"; do
  set -- $cfg
  echo "== HF_PT=$1 HF_MINB=$2"
  HF_PT=$1 HF_MINB=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-toys | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g main_ms %.3f frac %.3f share %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel_share_of_step']))"
done
