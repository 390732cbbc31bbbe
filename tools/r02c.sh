#!/bin/bash
# round-2 GPU call c (2 GPUs): parity of the fp32 landing path, its A/B, the 2-rank exchange test, bench at N=2 both exchanges
L=gpurun_out/r02c.log
python -m pytest tests/test_multirank.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4 > $L
for rep in 1 2; do
  python tools/kbench.py 4 13 0 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_noland13.so python tools/kbench.py 4 13 0 1 | tail -1 >> $L
  python tools/kbench.py 4 12 0 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_noland12.so python tools/kbench.py 4 12 0 1 | tail -1 >> $L
done
python tools/kbench.py 4 12 0 16 | tail -1 >> $L
SCN_LIB=scanner_b200/variants/lib_noland12.so python tools/kbench.py 4 12 0 16 | tail -1 >> $L
for x in peer nccl peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --exchange $x > gpurun_out/r02c_n2_$x.json 2>> gpurun_out/r02c_n2.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02c_n2_$x.json").read().strip().splitlines()[-1])
print("$x", "N=2 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), d["records_check"])
print("   step_ms", d["step_ms"])
PY
done
python bench.py --steps 40 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02c_n1.json 2>> gpurun_out/r02c_n2.err
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r02c_n1.json").read().strip().splitlines()[-1])
print("N=1 value", round(d["value"]/1e3,1), "Gs/s  ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4))
print("   step_ms", d["step_ms"])
PY
tail -5 gpurun_out/r02c_n2.err >> $L
cat $L
