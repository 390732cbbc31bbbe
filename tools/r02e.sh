#!/bin/bash
# round-2 GPU call e (1 GPU): new tests, wpt after the records hook, bench, ncu --set full of the p64 fp32 N=8192 (tile landing) and wpt kernels
L=gpurun_out/r02e.log; : > $L
python -m pytest tests/test_sweep_records.py tests/test_exchange.py tests/test_gpu_parity.py tests/test_abi.py tests/test_host_surface.py -m gpu -x -q 2>&1 | tail -6 >> $L
for rep in 1 2; do
  python tools/kbench.py 1 11 1 1 | tail -1 >> $L
  python tools/kbench.py 1 11 0 1 | tail -1 >> $L
done
python tools/kbench.py 3 12 0 64 | tail -1 >> $L
python tools/kbench.py 3 10 0 16 | tail -1 >> $L
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r02e_bench.json 2>> gpurun_out/r02e.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "e2e", round(d["e2e"]["value"]/1e3,2))
PY
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense -s 3 -c 1 -f -o gpurun_out/prof_r02e_p64_13 python tools/kbench.py 4 13 0 1 > gpurun_out/ncu_p64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense -s 3 -c 1 -f -o gpurun_out/prof_r02e_wpt python tools/kbench.py 1 11 1 1 > gpurun_out/ncu_wpt.log 2>&1
tail -3 gpurun_out/r02e.err >> $L
cat $L; ls -la gpurun_out | tail -5
