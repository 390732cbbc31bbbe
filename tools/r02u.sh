#!/bin/bash
L=gpurun_out/r02u.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
for rep in 1 2; do
  python tools/kbench.py 1 11 1 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_c4157.so python tools/kbench.py 1 11 1 1 | tail -1 >> $L
done
python tools/kbench.py 4 13 0 1 | tail -1 >> $L
python tools/kbench.py 3 12 0 64 | tail -1 >> $L
B=scanner_b200/scan_b200
$B bench 1 2048 8 1 4096 1500000 2 4096 1 1 0 | tail -1 >> $L
$B bench 1 2048 8 1 4096 3000000 2 8192 1 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 3 8192 6 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 4 8192 8 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 4 8192 10 64 200 | tail -1 >> $L
nproc >> $L
cat $L
