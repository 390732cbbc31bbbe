#!/bin/bash
# cluster kernel fp32: bulk L2 prefetch by one thread (default) vs per-lane prefetch.global.L2 (clpf1) vs none (clprev)
L=gpurun_out/r02zf.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -3 >> $L
for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "4 16 0 8" "4 14 0 4"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  for v in clpf1 clprev; do SCN_LIB=scanner_b200/variants/lib_$v.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L; done
done
cut -c1-110 $L
