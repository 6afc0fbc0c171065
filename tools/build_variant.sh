#!/bin/bash
# tools/build_variant.sh <tag> <nvcc -D flags...>: builds build/variants/libsvbfm_<tag>.so (same ABI, other kernel tuning macros)
# for A/B runs on the GPU box: SVBFM_LIB=build/variants/libsvbfm_<tag>.so python bench.py ...   (build/ is git-ignored, travels with gpurun)
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/scalable-variational-bayesian-factorization-machine_b200/csrc
out=$root/build/variants; mkdir -p $out/obj_$tag
ARCH="-gencode arch=compute_100a,code=sm_100a"
FL="-O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-O2,-Wno-unused-function -ccbin g++ --expt-relaxed-constexpr"
nvcc $FL "$@" -Xptxas -v -c $src/svbfm_engine.cu -o $out/obj_$tag/engine.o 2> $out/obj_$tag/ptxas.log || { tail -20 $out/obj_$tag/ptxas.log; exit 1; }
nvcc $FL "$@" -c $src/svbfm_ingest.cu -o $out/obj_$tag/ingest.o
nvcc $ARCH -shared -o $out/libsvbfm_$tag.so $out/obj_$tag/engine.o $out/obj_$tag/ingest.o -ldl -lcudart_static -lrt -lpthread
grep -A2 "k_streamILi1ELb1ELb1ELb1ELb0ELb1E" $out/obj_$tag/ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo " <- $tag"
