#!/bin/bash
# cluster kernel: window taps in shared memory + last twiddle moved to phase 2 + L2 prefetch (default build)
# vs the previous revision (variants/lib_clprev.so)
L=gpurun_out/r02zc.log; : > $L
timeout 600 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -5 >> $L
for rep in 1 2; do
  for cfg in "4 14 0 1" "4 15 0 1" "4 16 0 1" "1 15 1 1" "1 16 1 1" "2 16 1 4" "3 15 1 1" "4 16 0 8"; do
    timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
    SCN_LIB=scanner_b200/variants/lib_clprev.so timeout 120 python tools/kbench.py $cfg | tail -1 >> $L
  done
done
cat $L
