#!/bin/bash
# Build an experimental variant of libb200zk for A/B runs: tools/build_variant.sh NAME "-DFLAG ..." -> halo2-plonky2-verifier_b200/libb200zk_NAME.so
# (select it with B200ZK_LIB=.../libb200zk_NAME.so). Only the listed sources are rebuilt with the flags; the rest comes from build/.
set -e
NAME=$1; FLAGS=$2; shift 2
SRCS=${@:-ntt.cu}
cd "$(dirname "$0")/../halo2-plonky2-verifier_b200/csrc"
mkdir -p build_$NAME
OBJS=""
for f in *.cu; do
  o=build/${f%.cu}.o
  for s in $SRCS; do
    if [ "$s" = "$f" ]; then
      o=build_$NAME/${f%.cu}.o
      nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-pthread --expt-relaxed-constexpr $FLAGS -c $f -o $o
    fi
  done
  OBJS="$OBJS $o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libb200zk_$NAME.so $OBJS -Xlinker --version-script=exports.map
echo built ../libb200zk_$NAME.so
