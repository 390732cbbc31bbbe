#!/bin/bash
# headline kernel: no register prefetch of the next transform + one bulk L2 prefetch per warp (wptl2pf; wptl2pf5 = same
# at 5+ CTAs per SM) vs the default build (32 prefetch registers per lane)
L=gpurun_out/r02zm.log; : > $L
for rep in 1 2; do
  timeout 120 python tools/kbench.py 1 11 1 1 | tail -1 >> $L
  for v in wptl2pf wptl2pf5; do SCN_LIB=scanner_b200/variants/lib_$v.so timeout 120 python tools/kbench.py 1 11 1 1 | tail -1 >> $L; done
done
cut -c1-150 $L
