#!/bin/bash
# round-2 GPU call d (2 GPUs): bench at N=2 with both record exchanges, N=1 beside it
L=gpurun_out/r02d.log; : > $L
for x in peer nccl peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --exchange $x > gpurun_out/r02d_n2_$x.json 2>> gpurun_out/r02d.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02d_n2_$x.json").read().strip().splitlines()[-1])
print("$x", "N=2 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), d["records_check"])
print("   step_ms", d["step_ms"])
PY
done
python bench.py --steps 40 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02d_n1.json 2>> gpurun_out/r02d.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02d_n1.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --exchange peer > gpurun_out/r02d_n2_full.json 2>> gpurun_out/r02d.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02d_n2_full.json").read().strip().splitlines()[-1])
print("N=2 full: value", round(d["value"]/1e3,1), "e2e", d["e2e"])
PY
grep -i "error\|Traceback" gpurun_out/r02d.err | head -5 >> $L
cat $L
