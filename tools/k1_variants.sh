#!/bin/bash
# K1 tuning sweep (on the GPU box): average K1 launch time per variant (popc count x min blocks per SM)
for v in 53 54 63 64 43 44; do
  echo -n "variant $v: "; PLSTVO_K1_VARIANT=$v python bench.py --kernels-only 2>/dev/null | tail -1
done
