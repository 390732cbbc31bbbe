#!/bin/bash
# cluster kernel: dB rows staged in the dead S region and written by TMA bulk stores (default) vs 64 STG per thread (lib_prev)
L=gpurun_out/r02zk.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 >> $L
for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "1 14 1 1" "1 15 1 1" "1 16 1 1" "2 16 1 4" "3 15 1 1"; do
  timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_prev.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
done
cut -c1-118 $L
