#!/bin/bash
# P64 kernels (N = 4096 / 8192): bulk L2 prefetch of the next buffer (default) vs previous build (variants/lib_prev.so)
L=gpurun_out/r02zg.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3 >> $L
for rep in 1 2; do
for cfg in "4 13 0 1" "4 12 0 1" "2 12 1 64" "1 12 1 1" "1 13 1 1" "2 13 1 1"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_prev.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
done
done
cut -c1-110 $L
