#!/bin/bash
L=gpurun_out/r02o.log; : > $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 >> $L
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
tail -3 gpurun_out/r02o_bench.err >> $L
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02o_bench.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]/1e3,2), "ceiling", round(d["e2e"]["h2d_ceiling_gbs"],1))
print("sustained", round(d["sustained"]["msamples_per_s"]/1e3,1), d["sustained"]["clocks"])
print("plugin", d["plugin_e2e"])
for r in d["per_config"]: print(r["config"], r["kind"], r["fft_size"], r["averaging"], round(r["msamples_per_s"]/1e3,1), round(r["hbm_frac"],3), r["kernel"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
cat $L
