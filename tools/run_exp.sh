for rep in 1 2; do
python tools/kbench.py 4 13 0 1 | tail -1
SCN_LIB=scanner_b200/variants/lib_np13.so python tools/kbench.py 4 13 0 1 | tail -1
done
python tools/kbench.py 4 12 0 1 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
