#!/bin/bash
# cluster kernel: 2 rows per CTA (128 threads, 2 CTAs per SM, clusters of 2/4/8) vs 4 rows per CTA (default)
L=gpurun_out/r02za.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -3 >> $L
SCN_LIB=scanner_b200/variants/lib_clrows2.so timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -5 >> $L
for rep in 1 2; do
  for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "1 15 1 1" "1 16 1 1" "2 16 1 4"; do
    timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
    SCN_LIB=scanner_b200/variants/lib_clrows2.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  done
done
cat $L
