#!/bin/bash
# final tree after the headline-kernel change: parity + golden + full-size tests, then the default bench line
L=gpurun_out/r02zn.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $L
timeout 600 python bench.py > gpurun_out/r02zn_bench.json 2>> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r02zn_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["sustained"]["msamples_per_s"], d["parity"], d["plugin_e2e"]["value"])
PY
cat $L
