#!/bin/bash
# round-2 GPU call i (2 GPUs): SweepProcessor on real devices (NCCL in-process + peer windows), plugin-surface points
L=gpurun_out/r02i.log; : > $L
python -m pytest tests/test_host_surface.py tests/test_record.py tests/test_exchange.py -m gpu -x -q 2>&1 | tail -5 >> $L
B=scanner_b200/scan_b200
$B sweep 1 2048 20000000 8 1 20.0 2400000000.0 3150000000.0 64 3 7 2 1 nccl 1 2>&1 | grep -E "^sweep 1 step (0|24|25|49) |buffers" >> $L
$B sweep 1 2048 20000000 8 1 20.0 2400000000.0 3150000000.0 64 3 7 2 1 peer 1 2>&1 | grep -E "^sweep 1 step (0|24|25|49) |buffers" >> $L
$B bench 1 2048 8 1 4096 1500000 2 4096 1 1 0 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 3 8192 8 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 4 8192 8 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 3 8192 10 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 8000000 3 16384 8 128 300 | tail -1 >> $L
nproc >> $L
cat $L
