#!/bin/bash
# round-2 GPU call f (1 GPU): full GPU suite after the N=8192 pairwise-swap pass, kbench of the 8192 kinds, bench N=1 (side-stream summarize)
L=gpurun_out/r02f.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
for a in "4 13 0 1" "1 13 1 1" "3 13 1 1" "4 13 0 4" "3 13 0 8" "4 12 0 1" "1 11 1 1"; do
  python tools/kbench.py $a | tail -1 >> $L
done
python bench.py --no-extras --no-cpu-baseline --steps 40 > gpurun_out/r02f_bench.json 2>> gpurun_out/r02f.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "e2e", round(d["e2e"]["value"]/1e3,2))
print("   step_ms", d["step_ms"])
PY
tail -3 gpurun_out/r02f.err >> $L
cat $L
