#!/bin/bash
# development build of the fp32 library with only the German-credit kernels instantiated
# usage: ./build_dev.sh NAME [extra nvcc flags...]   ->  build_dev/libarp_NAME.so   (use with ARP_LIB_F32=...)
set -e
name=$1; shift
cd "$(dirname "$0")/autoreparam_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
  -DARP_DEV_GERMAN_ONLY "$@" arp_lib.cu -o ../../build_dev/libarp_$name.so
