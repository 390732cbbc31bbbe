#!/bin/bash
# A/B runs prepared at the end of round 1 (no GPU minutes were left to take them).  Build here, run under gpurun:
#   tools/next_round_ab.sh build      # on the build host: experiment libraries into scanner_b200/variants/
#   gpurun --timeout 600 -- 'bash tools/next_round_ab.sh run'
# Every line of output is one tools/kbench.py measurement (kernel-only, CUDA events, 768 MB of input).
set -e
if [ "$1" = build ]; then
  tools/build_variant.sh pairsync 11 -DSCN_WPT_PAIRSYNC=1          # WPT: warps of a CTA in lockstep (L0 I-cache)
  tools/build_variant.sh wpt5 11 -DSCN_WPT_MINCTAS=5               # WPT: 5 CTAs/SM (204 registers)
  tools/build_variant.sh pre13_0 13 -DSCN_P64_PREFETCH=0           # P64 fp32 N=8192 without the register prefetch
  tools/build_variant.sh pre13_8 13 -DSCN_P64_PREFETCH_13=8        # ... 8 points
  exit 0
fi
for rep in 1 2; do
  python tools/kbench.py 1 11 1 1 | tail -1
  SCN_LIB=scanner_b200/variants/lib_pairsync.so python tools/kbench.py 1 11 1 1 | tail -1
  SCN_LIB=scanner_b200/variants/lib_wpt5.so python tools/kbench.py 1 11 1 1 | tail -1
  python tools/kbench.py 4 13 0 1 | tail -1
  SCN_LIB=scanner_b200/variants/lib_pre13_0.so python tools/kbench.py 4 13 0 1 | tail -1
  SCN_LIB=scanner_b200/variants/lib_pre13_8.so python tools/kbench.py 4 13 0 1 | tail -1
done
# correctness of the lockstep experiment before believing its number
SCN_LIB=scanner_b200/variants/lib_pairsync.so python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cfg2 or (all_sizes and 11)" 2>&1 | tail -2
# plugin surface: staging copy vs zero-copy batches from the queue's slab (ProcessSamples::SetZeroCopy)
for w in 1 2; do
  scanner_b200/scan_b200 bench 1 2048 8 1 4096 2000000 $w | tail -1
  SCN_ZERO_COPY=1 scanner_b200/scan_b200 bench 1 2048 8 1 4096 2000000 $w | tail -1
done
scanner_b200/scan_b200 bench 4 8192 0 0 512 200000 2 | tail -1
SCN_ZERO_COPY=1 scanner_b200/scan_b200 bench 4 8192 0 0 512 200000 2 | tail -1
# and its correctness on the real library
SCN_ZERO_COPY=1 python -m pytest tests/test_host_surface.py tests/test_record.py -q -m gpu 2>&1 | tail -2
