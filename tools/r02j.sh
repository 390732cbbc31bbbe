#!/bin/bash
# round-2 GPU call j (1 GPU): first run of the cluster kernels (N = 2^14..2^16), the sweep-vs-synth text diff, wpt A/B
L=gpurun_out/r02j.log; : > $L
B=scanner_b200/scan_b200
A="1 1024 20000000 8 1 17.0 2400000000.0 2520000000.0 12 4 7"
TZ=UTC $B synth $A 1 1 2>/dev/null | grep -v -E "process thread|source thread|Frequency [0-9]+:|Elapsed" | sed 's/Start scan at .*/Start scan/' > /tmp/a.txt
TZ=UTC $B sweep $A 1 1 nccl 0 2>/tmp/err.txt | grep -v -E "process thread|source thread|Frequency [0-9]+:|Elapsed" | sed 's/Start scan at .*/Start scan/' > /tmp/b.txt
echo "synth lines $(wc -l < /tmp/a.txt) sweep lines $(wc -l < /tmp/b.txt)" >> $L
diff /tmp/a.txt /tmp/b.txt | head -20 >> $L
tail -2 /tmp/err.txt >> $L
for rep in 1 2; do
  python tools/kbench.py 1 11 1 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_wqagg.so python tools/kbench.py 1 11 1 1 | tail -1 >> $L
  SCN_LIB=scanner_b200/variants/lib_tw1.so python tools/kbench.py 1 11 1 1 | tail -1 >> $L
done
timeout 300 python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -15 >> $L
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "14 or 15 or 16" 2>&1 | tail -8 >> $L
for a in "4 14 0 1" "1 14 1 1" "4 15 0 1" "1 15 1 1" "4 16 0 1" "1 16 1 1" "3 16 0 4"; do
  timeout 120 python tools/kbench.py $a | tail -1 >> $L
  SCN_FOUR_STEP=1 timeout 120 python tools/kbench.py $a | tail -1 >> $L
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
cat $L
