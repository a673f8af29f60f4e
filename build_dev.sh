#!/bin/bash
# development builds of the fp32 library (kernel experiments; use with ARP_LIB_F32=build_dev/libarp_NAME.so)
# usage: ./build_dev.sh NAME [extra nvcc flags...]        German-credit kernels only (~1.5 min)
#        FULL=1 ./build_dev.sh NAME [extra nvcc flags...]  every model (~4 min)
set -e
name=$1; shift
cd "$(dirname "$0")/autoreparam_b200/csrc"
only=-DARP_DEV_GERMAN_ONLY
[ -n "$FULL" ] && only=
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
  $only "$@" arp_lib.cu -o ../../build_dev/libarp_$name.so
