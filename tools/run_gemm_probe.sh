#!/bin/bash
# Bring-up sweep for the tcgen05 GEMM; every config in its own process under a timeout so that a
# wedged kernel becomes a logged failure instead of a hung box.
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 90 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG
#    dtype passes ma mb   M     N    K   bn epi split iters
run  0 1 0 0   128   128    32  128 0 1 1
run  0 1 0 0   128   256    32  256 0 1 1
run  0 1 0 0   256   256   128  128 0 1 1
run  0 1 0 0  1024  1024  1024  256 0 1 5
run  0 1 0 1   128   128    32  128 0 1 1
run  0 1 0 1  1024  1024  1024  256 0 1 5
run  0 1 1 1   128   128    32  128 0 1 1
run  0 1 1 1  1024  1024  1024  256 0 1 5
run  0 1 1 0  1024  1024  1024  256 0 1 5
run  0 3 0 0   128   128    32  128 0 1 1
run  0 3 0 0  1024  1024  1024  128 0 1 5
run  0 3 0 0  1024  1024  1024  256 0 1 5
run  0 3 0 1  1024  1024  1024  128 0 1 5
run  0 3 1 1  1024  1024  1024  128 0 1 5
run  1 1 0 0  1024  1024  1024  256 0 1 5
run  1 1 0 1  1024  1024  1024  256 0 1 5
run  1 1 1 1  1024  1024  1024  256 0 1 5
# ragged + bias epilogue + split-K
run  0 1 0 0  1000   520   200  256 0 1 1
run  0 3 0 1   777   300   136  128 0 1 1
run  0 1 1 1   300   260  1000  256 0 1 1
run  0 1 0 0  1024  1024  1024  256 2 1 1
run  0 1 1 1  1024  1024 65536  256 1 0 5
run  0 3 1 1  1024  1024 65536  128 1 0 5
# config-2 shapes: fwd, dX, dW in both precisions, both tile widths
for bn in 128 256; do
run  0 1 0 0 65536  1024  1024  $bn 0 1 10
run  0 1 0 1 65536  1024  1024  $bn 0 1 10
run  0 1 1 1  1024  1024 65536  $bn 1 0 10
run  0 3 0 0 65536  1024  1024  $bn 0 1 5
run  0 3 0 1 65536  1024  1024  $bn 0 1 5
run  0 3 1 1  1024  1024 65536  $bn 1 0 5
done
# config-4 shape at 1/8 batch (bf16)
run  1 1 0 0 32768  4096  4096  256 0 1 5
run  1 1 0 1 32768  4096  4096  256 0 1 5
run  1 1 1 1  4096  4096 32768  256 1 0 5
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro" $LOG | tail -150
