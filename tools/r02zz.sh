#!/bin/bash
# final tree: full GPU suite, build()+smoke(), default bench line, reference arm
L=gpurun_out/r02zz.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
timeout 600 python bench.py > gpurun_out/r02zz_bench.json 2>> $L
tail -c 600 gpurun_out/r02zz_bench.json >> $L
cat $L
