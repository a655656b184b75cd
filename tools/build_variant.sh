#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   -> variants/<name>/libgwat_b200.so
# A kernel-experiment build of the engine translation unit (slim: 4 families x 2 detector counts, builds in about a minute)
# linked with the shipped objects of the other translation units.  Select it with GWAT_B200_LIB=$PWD/variants/<name>/libgwat_b200.so.
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/gw_analysis_tools_b200/csrc
out=$here/variants/$name
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 --diag-suppress 128 -Xptxas -v \
     -DGWAT_EXPERIMENT_SLIM "$@" -c -o $out/gwat_engine.o $src/gwat_engine.cu 2> $out/ptxas.log || { tail -30 $out/ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libgwat_b200.so $out/gwat_engine.o \
     $src/_obj/gwat_sampler.o $src/_obj/gwat_maximized.o $src/_obj/gwat_grids.o $src/_obj/gwat_losc.o $src/_obj/gwat_autocorr.o $src/_obj/gwat_queue.o $src/_obj/gwat_noise.o $src/_obj/gwat_chain_io.o -lcufft
echo built $out/libgwat_b200.so
