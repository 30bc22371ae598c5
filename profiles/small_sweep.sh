#!/bin/bash
# numfact time vs the batched-kernel threshold (fronts with s1+s2 <= HPDDM_B200_SMALL skip cuSOLVER), m = 128
for t in 160 256 384 512; do
  echo -n "SMALL=$t: "; HPDDM_B200_SMALL=$t python profiles/run_solve.py 128 1 2>&1 | grep -E "numfact|residual" | sed -e "s/{.*numfact_seconds/numfact_seconds/" | tr '\n' ' '; echo
done
