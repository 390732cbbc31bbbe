#!/bin/bash
# times a variant on the bench workload (with hits): prints value and kernel Msamples/s
python bench.py --no-e2e --no-cpu-baseline | python -c "import json,sys,os; d=json.loads(sys.stdin.readline()); print(os.path.basename(os.environ.get('SCN_LIB','default')), 'bench', round(d['value']/1e3,1), 'kernel', round(d['roofline']['kernel_msamples_per_s']/1e3,1), d['roofline']['kernel'], d['parity']['masks_equal'])"
