#!/bin/bash
# CTA-pair (cta_group::2) bring-up sweep; every config in its own process under a timeout.
mkdir -p gpurun_out
LOG=gpurun_out/probe_cg2.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 60 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
#    dtype passes ma mb   M     N    K   bn epi split iters ctas chunk notma cg
run  0 1 0 0   256   256    32  256 0 1 1 0 0 0 2
run  0 1 0 0   256   256   128  256 0 1 1 0 0 0 2
run  0 1 0 0  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 1 0 1  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 1 1 1  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 1 1 0  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 1 0 0  1024  1024  1024  128 0 1 3 0 0 0 2
run  1 1 0 0  1024  1024  1024  256 0 1 3 0 0 0 2
run  1 1 1 1  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 3 0 0   256   256    32  256 0 1 1 0 0 0 2
run  0 3 0 0  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 3 0 1  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 3 1 1  1024  1024  1024  256 0 1 3 0 0 0 2
run  0 3 0 0  1024  1024  1024  128 0 1 3 0 0 0 2
# ragged
run  0 1 0 0  1000   520   200  256 0 1 1 0 0 0 2
run  0 3 0 1   777   300   136  128 0 1 1 0 0 0 2
run  0 1 1 1   300   260  1000  256 0 1 1 0 0 0 2
run  0 1 0 0  1024  1024  1024  256 2 1 1 0 0 0 2
# config-2 shapes, both group sizes
for cg in 1 2; do
run  0 1 0 0 65536  1024  1024  256 0 1 10 0 0 0 $cg
run  0 1 0 1 65536  1024  1024  256 0 1 10 0 0 0 $cg
run  0 1 1 1  1024  1024 65536  256 1 0 10 0 0 0 $cg
run  0 3 0 0 65536  1024  1024  256 0 1 5 0 0 0 $cg
run  0 3 0 1 65536  1024  1024  256 0 1 5 0 0 0 $cg
run  0 3 1 1  1024  1024 65536  256 1 0 5 0 0 0 $cg
run  1 1 0 0 32768  4096  4096  256 0 1 5 0 0 0 $cg
run  1 1 0 1 32768  4096  4096  256 0 1 5 0 0 0 $cg
run  1 1 1 1  4096  4096 32768  256 1 0 5 0 0 0 $cg
done
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro" $LOG | tail -150
