#!/bin/bash
# Build A/B variants of libmfgp_b200.so that differ in -D tuning macros of ONE source file.
#   tools/build_variants.sh assemble.cu  name1 "-DMFGP_ASM_RI=2 -DMFGP_ASM_CTAS=3"  name2 "..."
# Output: tools/variants/libmfgp_<name>.so (git-ignored, travels with gpurun); use with MFGP_LIB=...
set -e
cd "$(dirname "$0")/.."
CSRC=multifidelity_datafusion_gps_b200/csrc
SRC=$1; shift
mkdir -p tools/variants
python multifidelity_datafusion_gps_b200/build.py > /dev/null
OTHERS=$(ls $CSRC/*.o | grep -v "${SRC%.cu}.o")
while [ $# -gt 0 ]; do
  NAME=$1; FLAGS=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $FLAGS \
       -c $CSRC/$SRC -o tools/variants/${SRC%.cu}_$NAME.o
  nvcc -shared -o tools/variants/libmfgp_$NAME.so tools/variants/${SRC%.cu}_$NAME.o $OTHERS -lcudart
  echo "built tools/variants/libmfgp_$NAME.so ($FLAGS)"
done
