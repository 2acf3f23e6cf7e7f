#!/bin/bash
# Timing-only knockout builds of the backward kernel (results are WRONG by construction): each -DKO_x removes one
# component (KO_S KO_DP KO_DV KO_DQ KO_DK: that MMA group; KO_EXP: the ex2; KO_DS: dS compute+store+fence; KO_FENCE: only
# the proxy fence; KO_DRAIN: the dQ drain; KO_RED: only its reductions; KO_TMA: Q/dO loads after the first ring fill).
# usage: tools/build_bwd_ko.sh <name> -DKO_x [...]   -> build/libaki_ko_<name>.so   (other objects from build/obj)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/obj_ko
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c $root/aki_b200/csrc/attn_bwd_sm100.cu -o $root/build/obj_ko/bwd_$name.o
o=$root/build/obj
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/build/libaki_ko_$name.so $o/api.o $o/meta.o $o/rope.o $o/decode.o \
  $o/attn_simt.o $o/attn_fwd_sm100.o $root/build/obj_ko/bwd_$name.o -cudart static
