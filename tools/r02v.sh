#!/bin/bash
L=gpurun_out/r02v.log; : > $L
run() {
  name=$1; g=$2; shift 2
  if [ $g -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g "$@" > gpurun_out/r02v_$name.json 2>> gpurun_out/r02v.err
  else
    python bench.py "$@" > gpurun_out/r02v_$name.json 2>> gpurun_out/r02v.err
  fi
  python - <<PY >> $L
import json
try:
    d=json.loads(open("gpurun_out/r02v_$name.json").read().strip().splitlines()[-1])
    sm=sorted(d["step_ms"])
    print("$name", "value", round(d["value"]/1e3,1), "Gs/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "step_ms min/med/max", sm[0], sm[len(sm)//2], sm[-1], (d.get("records_check") or {}).get("exchange"))
except Exception as e:
    print("$name FAILED", e)
PY
}
python -m pytest tests/test_exchange.py tests/test_multirank.py -m gpu -q 2>&1 | tail -2 >> $L
run n1 1 --steps 60 --no-e2e --no-extras --no-cpu-baseline
run n2_peer 2 --steps 60 --no-e2e --exchange peer
run n1_cfg4 1 --steps 60 --no-e2e --no-extras --no-cpu-baseline --workload cfg4
run n2_cfg4 2 --steps 60 --no-e2e --workload cfg4
run n1_cfg3 1 --steps 60 --no-e2e --no-extras --no-cpu-baseline --workload cfg3
run n2_cfg3 2 --steps 60 --no-e2e --workload cfg3
grep -i "error\|Traceback" gpurun_out/r02v.err | head -5 >> $L
cat $L
