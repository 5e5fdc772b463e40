#!/bin/bash
# K7 tuning sweep (models per thread x unroll): builds variants of ransac.cu on the GPU box and times the scoring kernel.
IFS=","; for V in ${SWEEP:-4 16,4 32,2 16,3 16,2 32,5 8}; do IFS=" "
  set -- $V
  (cd points_matching_b200/csrc && rm -f ransac.o && make -s EXTRA="-DSC_MPT_V=$1 -DSC_UNROLL_V=$2" 2>&1 | grep -E "error" ; grep -A1 "ransac_score_kernelILi0" ransac.ptxas.log | grep -o "Used [0-9]* registers" | head -1)
  python bench.py --steps 5 --warmup 3 --no-cpu --no-hamming --no-cfg5 --no-ramp --ransac-steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('MPT=$1 UNROLL=$2 k7_ms %.2f frac %.3f hyp/s %.3e' % (s['roofline']['kernel_ms'], s['roofline']['frac'], s['value']))"
done
