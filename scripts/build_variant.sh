#!/bin/bash
# scripts/build_variant.sh NAME [-DFLAG ...]: a differently-flagged build of the library for kernel
# experiments (build/libfrs_NAME.so, selected at run time with FRS_B200_LIB=...)
set -e
cd "$(dirname "$0")/../financial_rag_system_b200/csrc"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -shared \
  -o ../../build/libfrs_$name.so scan.cu index.cu exchange.cu bert.cu bert_fp32.cu encoder.cu
