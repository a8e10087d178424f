#!/bin/bash
# build the product library and the profiling variant; stop on any error
set -e
cd /root/repo/drake_ddp_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -lineinfo"
nvcc $F -Xptxas -v $EXTRA ddp_api.cu -o /tmp/v/lib_new.so 2> /tmp/v/build.log || { grep -m5 error /tmp/v/build.log; exit 1; }
grep -A2 "${KERNEL_GREP:-backward_mma_kernelINS_9QuadrupedEEE}" /tmp/v/build.log | grep -v "^--" | grep -v Compiling
nvcc $F -DDDP_BWD_PROFILE $EXTRA ddp_api.cu -o /root/repo/scratch/out/lib_prof.so
cp /tmp/v/lib_new.so /root/repo/drake_ddp_b200/libddp_b200.so
echo BUILD_OK
