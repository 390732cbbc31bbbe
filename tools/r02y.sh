#!/bin/bash
# 8-GPU re-run after the fused exchange step: N=1, N=8 peer / nccl, host enqueue time per step
L=gpurun_out/r02y.log; : > $L
run() {
  name=$1; g=$2; shift 2
  if [ $g -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g "$@" > gpurun_out/r02y_$name.json 2>> gpurun_out/r02y.err
  else
    python bench.py "$@" > gpurun_out/r02y_$name.json 2>> gpurun_out/r02y.err
  fi
  python - <<PY >> $L
import json
try:
    d=json.loads(open("gpurun_out/r02y_$name.json").read().strip().splitlines()[-1])
    sm=sorted(d["step_ms"])
    print("$name", "value", round(d["value"]/1e3,1), "Gs/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "step_ms min/med/max", sm[0], sm[len(sm)//2], sm[-1], "host_enqueue_ms", round(d["host_enqueue_ms_per_step"],4), (d.get("records_check") or {}).get("exchange"))
except Exception as e:
    print("$name FAILED", e)
PY
}
python -m pytest tests/test_exchange.py -m gpu -q 2>&1 | tail -2 >> $L
run n1 1 --steps 60 --no-e2e --no-extras --no-cpu-baseline
run n8_peer 8 --steps 60 --no-e2e --exchange peer
run n8_nccl 8 --steps 60 --no-e2e --exchange nccl
run n4_peer 4 --steps 60 --no-e2e --exchange peer
run n2_peer 2 --steps 60 --no-e2e --exchange peer
run n8_cfg4 8 --steps 60 --no-e2e --workload cfg4
run n1_cfg4 1 --steps 60 --no-e2e --no-extras --no-cpu-baseline --workload cfg4
run n8_full 8 --steps 20
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02y_n8_full.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("n8_full value", round(d["value"]/1e3,1), "e2e", round(e["value"]/1e3,1), "h2d_gbs", round(e["h2d_gbs"],1), "ceiling", round(e["h2d_ceiling_gbs"],1))
PY
grep -i "error\|Traceback" gpurun_out/r02y.err | head -5 >> $L
cat $L
