#!/bin/bash
# round-2 GPU call h (2 GPUs): plugin-surface sweep with the ring slab, bench N=1 / N=2 (peer, nccl) with the side-stream records
L=gpurun_out/r02h.log; : > $L
python -m pytest tests/test_host_surface.py tests/test_record.py tests/test_hackrf_sweep.py tests/test_multirank.py -m gpu -x -q 2>&1 | tail -3 >> $L
B=scanner_b200/scan_b200
$B bench 1 2048 8 1 4096 1500000 2 4096 1 1 0 | tail -1 >> $L
$B bench 1 2048 8 1 4096 3000000 2 8192 1 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 6000000 2 8192 2 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 2 8192 6 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 3 8192 8 64 200 | tail -1 >> $L
SCN_STAGING_COPY=1 $B bench 1 2048 8 1 4096 8000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 4 8192 0 0 512 600000 2 2048 4 16 200 | tail -1 >> $L
nproc >> $L; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> $L
for x in peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --exchange $x > gpurun_out/r02h_n2_$x.json 2>> gpurun_out/r02h.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02h_n2_$x.json").read().strip().splitlines()[-1])
print("$x", "N=2 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), d["records_check"])
print("   step_ms", d["step_ms"])
PY
done
python bench.py --steps 40 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02h_n1.json 2>> gpurun_out/r02h.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02h_n1.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4))
PY
grep -i "error\|Traceback" gpurun_out/r02h.err | head -5 >> $L
cat $L
