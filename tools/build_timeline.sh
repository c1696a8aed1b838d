#!/bin/bash
# the -DS1S2_TIMELINE build of the library for tools/timeline.py (kept out of the package directory: never the shipped library)
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_libs
cd s1-to-s2_super-resolution_project-code_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -DS1S2_TIMELINE -o ../../ab_libs/lib_timeline.so s1s2_lib.cu
