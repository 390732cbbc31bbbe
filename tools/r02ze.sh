#!/bin/bash
# cluster kernel: L2 prefetch of the fp32 loads that are not register-prefetched (default) vs previous (variants/lib_clprev.so)
L=gpurun_out/r02ze.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -3 >> $L
for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "4 16 0 8" "4 14 0 4" "1 16 1 1"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_clprev.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
done
cut -c1-110 $L
