#!/bin/bash
# ncu of the cluster kernel after the mbarrier change, summarised ON the box
for cfg in "4 16 0 1 cl16_fp32" "4 14 0 1 cl14_fp32"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:spectrum_sense_cluster -s 2 -c 1 -f -o /tmp/prof_$5 python tools/kbench.py $1 $2 $3 $4 > /tmp/ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$5.ncu-rep > gpurun_out/r02zj_ncu_$5.txt 2>&1
  ncu -i /tmp/prof_$5.ncu-rep --page source --csv > /tmp/src_$5.csv 2>/dev/null
  python tools/ncu_stalls.py /tmp/src_$5.csv >> gpurun_out/r02zj_ncu_$5.txt 2>&1
done
ls -la gpurun_out | tail -3
