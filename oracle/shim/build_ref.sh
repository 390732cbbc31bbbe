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
# signalSource.cpp:36-38 (Start/Stop) and hackRFSource.cpp:150-178 (set_sample_rate) are non-void functions
# without a return statement: at -O2 gcc drops the epilogue and control runs on into the next function
# (observed: the constructor never returns), and gcc 13's -O0 default plants a trap there instead.
# -O0 -fno-unreachable-traps keeps the plain epilogue; the only arithmetic in these two files is the integer
# header parse / sample patch of hackRFSource.cpp:186-222, which -O0 does not change.
$CXX $FLAGS -O0 -fno-unreachable-traps -c "$REF/signalSource.cpp" -o "$OUT/signalSource.o"
$CXX $FLAGS -O0 -fno-unreachable-traps -c "$REF/hackRFSource.cpp" -o "$OUT/hackRFSource.o"
$CXX $FLAGS -c "$HERE/shim_impl.cpp" -o "$OUT/shim_impl.o"
$CXX $FLAGS -c "$HERE/shim_hackrf.cpp" -o "$OUT/shim_hackrf.o"
$CXX $FLAGS -c "$HERE/ref_tool.cpp" -o "$OUT/ref_tool.o"
$CXX -pthread -o "$OUT/ref_tool" "$OUT/ref_tool.o" "$OUT/utility.o" "$OUT/fft.o" "$OUT/process.o" "$OUT/frequencyTable.o" "$OUT/signalSource.o" "$OUT/hackRFSource.o" "$OUT/shim_impl.o" "$OUT/shim_hackrf.o"
$CXX -pthread -o "$OUT/ref_tool_cmath" "$OUT/ref_tool.o" "$OUT/utility_cmath.o" "$OUT/fft.o" "$OUT/process.o" "$OUT/frequencyTable.o" "$OUT/signalSource.o" "$OUT/hackRFSource.o" "$OUT/shim_impl.o" "$OUT/shim_hackrf.o"
rm -f "$OUT"/*.o
echo "oracle/_ref: built ref_tool, ref_tool_cmath from $REF"
