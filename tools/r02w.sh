#!/bin/bash
L=gpurun_out/r02w.log; : > $L
for rep in 1 2; do
  python tools/kbench.py 4 16 0 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_clpre24.so python tools/kbench.py 4 16 0 1 | tail -1 >> $L
  python tools/kbench.py 4 14 0 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_clpre24_14.so python tools/kbench.py 4 14 0 1 | tail -1 >> $L
done
cat $L
