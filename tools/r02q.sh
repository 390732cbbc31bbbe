#!/bin/bash
# round-2 8-GPU call: weak scaling of the headline workload (peer windows / NCCL) and of cfg4, N=1 beside it on the same box
L=gpurun_out/r02q.log; : > $L
run() {  # name gpus args...
  name=$1; g=$2; shift 2
  if [ $g -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g "$@" > gpurun_out/r02q_$name.json 2>> gpurun_out/r02q.err
  else
    python bench.py "$@" > gpurun_out/r02q_$name.json 2>> gpurun_out/r02q.err
  fi
  python - <<PY >> $L
import json
try:
    d=json.loads(open("gpurun_out/r02q_$name.json").read().strip().splitlines()[-1])
    sm=sorted(d["step_ms"])
    print("$name", "value", round(d["value"]/1e3,1), "Gs/s ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "step_ms min/med/max", sm[0], sm[len(sm)//2], sm[-1], d.get("records_check"))
except Exception as e:
    print("$name FAILED", e)
PY
}
run n1 1 --steps 40 --no-e2e --no-extras --no-cpu-baseline
run n8_peer 8 --steps 40 --no-e2e --exchange peer
run n8_nccl 8 --steps 40 --no-e2e --exchange nccl
run n4_peer 4 --steps 40 --no-e2e --exchange peer
run n2_peer 2 --steps 40 --no-e2e --exchange peer
run n1_cfg4 1 --steps 40 --no-e2e --no-extras --no-cpu-baseline --workload cfg4
run n8_cfg4 8 --steps 40 --no-e2e --workload cfg4
run n8_full 8 --steps 20
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02q_n8_full.json").read().strip().splitlines()[-1])
print("n8_full e2e", d["e2e"])
PY
grep -i "error\|Traceback" gpurun_out/r02q.err | head -5 >> $L
cat $L
