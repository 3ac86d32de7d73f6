#!/bin/bash
# TF32 + 2x bf16-correction mode (passes = 2): accuracy and speed vs 3xTF32, all operand layouts
mkdir -p gpurun_out
LOG=gpurun_out/probe_p2.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 60 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
#    dtype passes ma mb   M     N    K   bn epi split iters ctas chunk notma cg
for lay in "0 0" "0 1" "1 0" "1 1"; do
run  0 2 $lay   128   128    32  128 0 1 1 0 0 0 1
run  0 2 $lay  1024  1024  1024  256 0 1 3 0 0 0 1
run  0 2 $lay  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 2 $lay  1024  1024  1024  128 0 1 3 0 0 0 2
run  0 2 $lay  1000   520   200  256 0 1 1 0 0 0 2
done
run  0 2 0 1   777   300   136  128 0 1 1 0 0 0 2
run  0 2 1 1   300   260  1000  256 1 0 1 0 0 0 2
for p in 2 3; do
run  0 $p 0 0 65536  1024  1024  256 0 1 10 0 4 0 2
run  0 $p 0 1 65536  1024  1024  256 0 1 10 0 4 0 2
run  0 $p 1 1  1024  1024 65536  256 1 0 10 0 4 0 2
done
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro" $LOG | tail -80
