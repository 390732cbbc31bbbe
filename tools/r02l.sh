#!/bin/bash
L=gpurun_out/r02l.log; : > $L
for a in "1 14 0 1" "4 14 0 1" "4 15 0 1" "4 16 0 1" "3 16 1 2" "2 15 1 1" "1 16 1 1"; do
  echo "== $a" >> $L
  timeout 200 python tests/cluster_debug.py $a 2>&1 | tail -4 >> $L
done
timeout 400 python -m pytest tests/test_gpu_large.py -m gpu -q 2>&1 | tail -8 >> $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "14" 2>&1 | tail -5 >> $L
for a in "4 14 0 1" "1 14 1 1" "4 15 0 1" "1 15 1 1" "4 16 0 1" "1 16 1 1" "3 16 0 4"; do
  timeout 120 python tools/kbench.py $a | tail -1 >> $L
done
timeout 300 python -m pytest tests/test_host_surface.py -m gpu -q 2>&1 | tail -3 >> $L
cat $L
