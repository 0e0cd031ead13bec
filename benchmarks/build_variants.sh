#!/bin/bash
# A/B builds of ONE kernel file with a different -D knob, linked against the other (current) objects:
#   benchmarks/build_variants.sh sim_umma COR_SIM_POLY_OF4 0 1 3   ->  cor_b200/build/ab/libcor_b200_COR_SIM_POLY_OF4_<v>.so
# Select one at run time with COR_B200_LIB=<path>.  cor_b200/build/ is git-ignored but travels to the GPU box.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
FILE=$1; KNOB=$2; shift 2
python -m cor_b200.build > /dev/null
mkdir -p "$ROOT/cor_b200/build/ab"
for v in "$@"; do
  obj="$ROOT/cor_b200/build/ab/${FILE}_${KNOB}_${v}.o"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr \
       -I "$ROOT/include" -D${KNOB}=${v} -c "$ROOT/cor_b200/csrc/${FILE}.cu" -o "$obj"
  others=$(ls "$ROOT"/cor_b200/build/*.o | grep -v "/${FILE}.o")
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o "$ROOT/cor_b200/build/ab/libcor_b200_${KNOB}_${v}.so" $obj $others
  echo "built cor_b200/build/ab/libcor_b200_${KNOB}_${v}.so"
done
