#!/bin/bash
# cluster kernel, separated: default = window in smem + twiddle in phase 2; clwin = window only; cltw = twiddle only; clprev = neither
L=gpurun_out/r02zd.log; : > $L
for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "1 16 1 1" "2 16 1 4" "4 16 0 8"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  for v in clwin cltw clprev; do
    SCN_LIB=scanner_b200/variants/lib_$v.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  done
done
cut -c1-100 $L
