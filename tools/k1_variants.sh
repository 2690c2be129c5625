#!/bin/bash
# K1 tuning sweep (on the GPU box): average K1 launch time per variant
# (PLSTVO_K1_VARIANT = <queries per thread><popc per distance><min blocks per SM>)
for v in ${@:-253 254 263 243 452 453 462 442}; do
  echo -n "variant $v: "; PLSTVO_K1_VARIANT=$v python bench.py --kernels-only 2>/dev/null | tail -1
done
