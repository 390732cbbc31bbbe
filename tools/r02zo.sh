#!/bin/bash
# cluster kernel, integer kinds: 32 (default) / 16 / 0 of the 32 raw loads per thread held in registers one buffer ahead,
# the rest bulk-prefetched into L2
L=gpurun_out/r02zo.log; : > $L
for cfg in "1 14 1 1" "1 15 1 1" "1 16 1 1" "3 15 1 1" "3 16 1 4"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  for v in clpi16 clpi0; do SCN_LIB=scanner_b200/variants/lib_$v.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L; done
done
cut -c1-118 $L
