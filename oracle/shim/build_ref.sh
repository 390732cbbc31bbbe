#!/bin/sh
# Compiles the reference's own hot-path sources, where they lie (read-only), against the shim
# headers into oracle/_ref/.  Nothing from the reference tree is copied into this repo.
#   ref_tool        utility.cpp built with <math.h> visible  (log2/sqrt -> float overloads; GCC >= 6)
#   ref_tool_cmath  utility.cpp built with only <cmath>      (log2/sqrt -> double C functions)
# (utility.cpp includes neither header itself -- SURVEY.md section 0 defect 3 -- so the include
#  is forced on the command line; the two binaries pin both readings of its dB arithmetic.)
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref"
mkdir -p "$OUT"
CXX=${CXX:-g++}
# the reference's own flags are "-g -O3 -std=gnu++11" (Makefile:23); no -march, no fast-math
FLAGS="-O2 -std=gnu++11 -w -pthread -I$HERE -I$REF"
for v in math cmath; do
  if [ "$v" = math ]; then INC="-include math.h"; SUF=""; else INC="-include cmath"; SUF="_cmath"; fi
  $CXX $FLAGS $INC -c "$REF/utility.cpp" -o "$OUT/utility$SUF.o"
done
$CXX $FLAGS -c "$REF/fft.cpp" -o "$OUT/fft.o"
$CXX $FLAGS -c "$REF/process.cpp" -o "$OUT/process.o"
$CXX $FLAGS -c "$REF/frequencyTable.cpp" -o "$OUT/frequencyTable.o"
$CXX $FLAGS -c "$HERE/shim_impl.cpp" -o "$OUT/shim_impl.o"
$CXX $FLAGS -c "$HERE/ref_tool.cpp" -o "$OUT/ref_tool.o"
$CXX -pthread -o "$OUT/ref_tool" "$OUT/ref_tool.o" "$OUT/utility.o" "$OUT/fft.o" "$OUT/process.o" "$OUT/frequencyTable.o" "$OUT/shim_impl.o"
$CXX -pthread -o "$OUT/ref_tool_cmath" "$OUT/ref_tool.o" "$OUT/utility_cmath.o" "$OUT/fft.o" "$OUT/process.o" "$OUT/frequencyTable.o" "$OUT/shim_impl.o"
rm -f "$OUT"/*.o
echo "oracle/_ref: built ref_tool, ref_tool_cmath from $REF"
