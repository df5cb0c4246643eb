#!/bin/bash
# usage: build_variant.sh NAME "-DFOO=1 ..."   -> tools/variants/NAME.so
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $2 -shared -o /root/repo/tools/variants/$1.so /root/repo/chomp_b200/csrc/chomp_b200.cu
