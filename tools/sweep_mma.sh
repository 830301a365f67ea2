#!/bin/bash
for m in 0 1; do
  echo "== HF_MMA=$m"
  HF_MMA=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-toys | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g main_ms %.3f frac %.3f share %.3f ms/step %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['ms_per_step']))"
done
