#!/bin/bash
# cluster kernel: mask/prefix aliased onto the dead S region (smem under the 132 KB carve-out -> 124 KB L1) vs previous;
# and N = 2^14 int8: cluster kernel (SCN_CLUSTER_INT8_14=1) vs the 16-point kernel
L=gpurun_out/r02zh.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 >> $L
for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "1 15 1 1" "1 16 1 1" "2 16 1 4" "3 14 1 1"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_prev.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
done
timeout 120 python tools/kbench.py 1 14 1 1 | tail -1 >> $L
SCN_CLUSTER_INT8_14=1 timeout 120 python tools/kbench.py 1 14 1 1 | tail -1 >> $L
SCN_CLUSTER_INT8_14=1 SCN_LIB=scanner_b200/variants/lib_prev.so timeout 120 python tools/kbench.py 1 14 1 1 | tail -1 >> $L
timeout 120 python tools/kbench.py 1 14 0 1 | tail -1 >> $L
SCN_CLUSTER_INT8_14=1 timeout 120 python tools/kbench.py 1 14 0 1 | tail -1 >> $L
cut -c1-118 $L
