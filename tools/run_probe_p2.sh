#!/bin/bash
# TF32 + 2x bf16-correction mode (passes = 2), K-major operands: accuracy and speed vs 3xTF32
mkdir -p gpurun_out
LOG=gpurun_out/probe_p2.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 60 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
#    dtype passes ma mb   M     N    K   bn epi split iters ctas chunk notma cg
run  0 2 0 0   128   128    32  128 0 1 1 0 0 0 1
run  0 2 0 0   256   256    64  256 0 1 1 0 0 0 2
run  0 2 0 0  1024  1024  1024  256 0 1 3 0 0 0 1
run  0 2 0 0  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 2 0 0  1024  1024  1024  128 0 1 3 0 0 0 2
run  0 2 0 0  1000   520   200  256 0 1 1 0 0 0 2
run  0 2 0 0   777   300   136  128 0 1 1 0 0 0 2
run  0 2 0 0  1024  1024 65536  256 1 0 3 0 0 0 2
for ch in 4 8; do
run  0 2 0 0 65536  1024  1024  256 0 1 10 0 $ch 0 2
run  0 3 0 0 65536  1024  1024  256 0 1 10 0 $ch 0 2
done
run  0 2 0 0 65536  1024  1024  256 2 1 10 0 4 0 2
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro" $LOG | tail -60
