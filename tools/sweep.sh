#!/bin/bash
# BASELINE.json configs[4]: kernel-only throughput over FFT size x sample kind (tools/kbench.py, CUDA events,
# 768 MB of device-resident input per point, dB spectrum written).
for kind in 1 3 4; do
  for l in 8 9 10 11 12 13 14 15 16; do
    dc=0; [ $kind -ne 4 ] && dc=1
    python tools/kbench.py $kind $l $dc 1 2>&1 | tail -1
  done
done
python tools/kbench.py 3 10 0 16; python tools/kbench.py 3 12 0 64; python tools/kbench.py 1 11 1 16
