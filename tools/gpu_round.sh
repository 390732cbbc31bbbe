#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, bench, ncu launch list and one full capture.
# Usage: tools/gpu_round.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_$TAG.log
python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spectrum_sense|summarize_steps|merge_records|time_domain" -c 40 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/ncu_bench_$TAG.log 2>&1
# full capture of the fused kernel
ncu --set full --clock-control none --import-source on -k regex:spectrum_sense -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -12
