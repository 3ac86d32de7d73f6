#!/bin/bash
# 3xTF32 chunked-promotion sweep: accuracy and speed vs chunk size (k-blocks of 32 K-elements per TMEM chunk).
mkdir -p gpurun_out
LOG=gpurun_out/probe3.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 90 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
#    dtype passes ma mb   M     N    K   bn epi split iters ctas chunk
run  0 3 0 0   128   128    32  128 0 1 1 0 4
run  0 3 0 0   777   300   136  128 0 1 1 0 2
run  0 3 0 1  1000   520   200  256 0 1 1 0 4
run  0 3 1 1   300   260  1000  256 1 0 1 0 4
for ch in 1 2 4 8 16; do
for bn in 128 256; do
run  0 3 0 0 65536  1024  1024  $bn 0 1 5 0 $ch
done
done
for ch in 2 4 8; do
run  0 3 0 1 65536  1024  1024  256 0 1 5 0 $ch
run  0 3 1 1  1024  1024 65536  256 1 0 5 0 $ch
run  0 3 0 0 65536  1024  1024  256 2 1 5 0 $ch
done
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro" $LOG | tail -150
