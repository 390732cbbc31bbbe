for rep in 1 2; do
for cfg in "4 13 0" "1 13 1" "4 12 0" "4 11 0" "1 11 1"; do
set -- $cfg
python tools/kbench.py $1 $2 $3 1 | tail -1
SCN_LIB=scanner_b200/variants/lib_ns$2.so python tools/kbench.py $1 $2 $3 1 | tail -1
done; done
