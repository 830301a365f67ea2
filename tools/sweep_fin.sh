#!/bin/bash
# hf_finish_grad_kernel variants (points per thread x entries side by side) on the C4 bench step (tuning aid)
# build: HF_NVCC_EXTRA="-DHF_FIN_PPT=p -DHF_FIN_UNROLL=u" python -m pyhf_b200.build --force; cp pyhf_b200/libhf_b200.so variants/fin_p${p}_u${u}.so
cp pyhf_b200/libhf_b200.so /tmp/keep.so
for v in variants/*.so; do
  cp $v pyhf_b200/libhf_b200.so
  echo "== $v"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-toys | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms_per_step %.3f e2e %.4g main_ms %.3f share %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['kernel_share_of_step']))"
done
cp /tmp/keep.so pyhf_b200/libhf_b200.so
