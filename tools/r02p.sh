#!/bin/bash
# A/B of today's head against this morning's revision (4157fa3), same box, same call
L=gpurun_out/r02p.log; : > $L
for a in "1 12 1 1" "3 12 0 64" "4 12 0 1" "1 13 1 1" "4 13 0 1" "3 12 1 1"; do
  python tools/kbench.py $a | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_c4157.so python tools/kbench.py $a | tail -1 >> $L
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 >> $L
cat $L
