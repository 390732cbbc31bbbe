#!/bin/bash
L=gpurun_out/r02k.log; : > $L
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tests/cluster_debug.py 1 14 0 1 2>&1 | grep -v "^=========     Host Frame\|^=========         in\|^=========         \*" | head -60 >> $L
echo ---- >> $L
timeout 120 python tests/cluster_debug.py 4 14 0 1 2>&1 | tail -4 >> $L
timeout 120 python tests/cluster_debug.py 4 15 0 1 2>&1 | tail -4 >> $L
timeout 120 python tests/cluster_debug.py 4 16 0 1 2>&1 | tail -4 >> $L
timeout 120 python tests/cluster_debug.py 3 16 1 2 2>&1 | tail -4 >> $L
cat $L
