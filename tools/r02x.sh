#!/bin/bash
L=gpurun_out/r02x.log; : > $L
python -m pytest tests/test_host_surface.py -m gpu -q 2>&1 | tail -2 >> $L
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense_wpt -s 3 -c 1 -f -o /tmp/prof_wpt python tools/kbench.py 1 11 1 1 > /tmp/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_wpt.ncu-rep > gpurun_out/r02x_ncu_wpt.txt 2>&1
ncu -i /tmp/prof_wpt.ncu-rep --page source --csv > /tmp/src_wpt.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/src_wpt.csv >> gpurun_out/r02x_ncu_wpt.txt 2>&1
cat $L
