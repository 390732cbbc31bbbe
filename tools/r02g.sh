#!/bin/bash
# round-2 GPU call g (1 GPU): host-surface GPU tests after the ingest rework, plugin-surface sweep, ncu of the N=8192 fp32 kernel
L=gpurun_out/r02g.log; : > $L
python -m pytest tests/test_host_surface.py tests/test_record.py tests/test_hackrf_sweep.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4 >> $L
B=scanner_b200/scan_b200
# kind N enob dc distinct total workers max_batch producers append linger
$B bench 1 2048 8 1 4096 1500000 2 4096 1 1 0 | tail -1 >> $L
SCN_STAGING_COPY=1 $B bench 1 2048 8 1 4096 1500000 2 4096 1 1 0 | tail -1 >> $L
$B bench 1 2048 8 1 4096 3000000 2 8192 1 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 6000000 2 8192 2 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 6000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 6000000 3 8192 6 64 200 | tail -1 >> $L
$B bench 1 2048 8 1 4096 6000000 2 16384 4 256 300 | tail -1 >> $L
SCN_STAGING_COPY=1 $B bench 1 2048 8 1 4096 6000000 2 8192 4 64 200 | tail -1 >> $L
$B bench 4 8192 0 0 512 400000 2 2048 4 16 200 | tail -1 >> $L
nproc >> $L
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense -s 3 -c 1 -f -o gpurun_out/prof_r02g_p64_13 python tools/kbench.py 4 13 0 1 > gpurun_out/ncu_p64g.log 2>&1
cat $L
