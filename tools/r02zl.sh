#!/bin/bash
# ncu of the cfg4 kernel (P64 fp32 N = 8192) and the headline kernel with the per-section stall breakdown, summarised ON the box
for cfg in "4 13 0 1 p64_fp32_8192 spectrum_sense_p64" "1 11 1 1 wpt_int8_2048 spectrum_sense_wpt"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$6 -s 2 -c 1 -f -o /tmp/prof_$5 python tools/kbench.py $1 $2 $3 $4 > /tmp/ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$5.ncu-rep > gpurun_out/r02zl_ncu_$5.txt 2>&1
  ncu -i /tmp/prof_$5.ncu-rep --page source --csv > /tmp/src_$5.csv 2>/dev/null
  python tools/ncu_stalls.py /tmp/src_$5.csv >> gpurun_out/r02zl_ncu_$5.txt 2>&1
done
ls -la gpurun_out | tail -3
