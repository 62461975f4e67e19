#!/bin/bash
# developer A/B helper: build a library variant  tools/build_variant.sh <tag> [-DFLAG ...]  ->  csrc/ab_<tag>.so
# (loaded through FSD_LIBFSDPLAN, see tools/mode_ab.py)
set -e
tag=$1; shift
C=$(dirname "$0")/../ft_fsd_path_planning_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xcompiler -Wno-unknown-pragmas \
  "$@" -o $C/ab_$tag.so $C/kernels.cu $C/kernels_big.cu $C/cpu_backend.cpp
echo built $C/ab_$tag.so
