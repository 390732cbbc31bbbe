#!/bin/bash
# 1 GPU: group-max detection A/B through bench.py (real hit rate), launch list of the bench command, full capture of the headline kernel
L=gpurun_out/r02t.log; : > $L
for rep in 1 2; do
  for lib in "" scanner_b200/variants/lib_gmax.so; do
    SCN_LIB=$lib python bench.py --steps 40 --no-e2e --no-extras --no-cpu-baseline > /tmp/b.json 2>> gpurun_out/r02t.err
    python - <<PY >> $L
import json
d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("${lib:-default}", "value", round(d["value"]/1e3,1), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "parity", d["parity"])
PY
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spectrum_sense|summarize_steps|merge_records|publish|exchange" -c 60 --csv --log-file gpurun_out/r02t_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /tmp/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense_wpt -s 4 -c 1 -f -o /tmp/prof_wpt python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /tmp/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_wpt.ncu-rep > gpurun_out/r02t_ncu_wpt.txt 2>&1
ncu -i /tmp/prof_wpt.ncu-rep --page source --csv > /tmp/src_wpt.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/src_wpt.csv >> gpurun_out/r02t_ncu_wpt.txt 2>&1
tail -3 gpurun_out/r02t.err >> $L
cat $L; tail -5 gpurun_out/r02t_launches.csv | cut -c1-200
