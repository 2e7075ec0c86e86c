#!/bin/bash
# experiment builds: tools/build_alt.sh NAME [-DFLAG=...]...  -> fastenhancer_b200/_alt/NAME.so (only the 16 kHz B / T kernels are
# recompiled with the flags; everything else links from the default build).  Select at run time with FE_LIB=...
set -e
cd "$(dirname "$0")/../fastenhancer_b200"
NAME=$1; shift
mkdir -p _alt/obj_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
for u in ${UNITS:-fe_inst_16b fe_inst_16t}; do nvcc $FLAGS "$@" -c csrc/$u.cu -o _alt/obj_$NAME/$u.o & done; wait
OBJS=$(ls _build/*.o | grep -v -E "$(echo ${UNITS:-fe_inst_16b fe_inst_16t} | tr " " "|")")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o _alt/$NAME.so $OBJS _alt/obj_$NAME/*.o
echo built _alt/$NAME.so
