#!/bin/bash
# F16X3 (fp16 pair, 3 passes) bring-up + tuning sweep; every config in its own process under a timeout.
mkdir -p gpurun_out
LOG=gpurun_out/probe_f16x3.log
: > $LOG
P=tools/gemm_probe
run() { echo "== $*" >> $LOG; timeout 90 $P "$@" >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG
#    dtype passes ma mb   M     N    K   bn epi split iters ctas chunk notma cg
run  2 0 0 0   128   128    64  128 0 1 1 0 0 0 1
run  2 0 0 0   128   256    64  256 0 1 1 0 0 0 1
run  2 0 0 0   256   256   128  256 0 1 1 0 0 0 2
run  2 0 0 0  1024  1024  1024  256 0 1 5
run  2 0 0 1  1024  1024  1024  256 0 1 5
run  2 0 1 1  1024  1024  1024  256 0 1 5
run  2 0 1 0  1024  1024  1024  256 0 1 5
run  2 0 0 0  1024  1024  1024  128 0 1 5
run  2 0 0 0  1024  1024  1024  256 0 1 5 0 0 0 1
run  2 0 0 0  1000   520   200  256 0 1 1
run  2 0 0 1   777   304   136  128 0 1 1
run  2 0 1 1   304   264  1000  256 0 1 1
run  2 0 0 0  1024  1024  1024  256 2 1 1
run  2 0 1 1  1024  1024  8192  256 1 0 5
# config-2 shapes, chunk sweep
for ch in 1 2 4 8; do
run  2 0 0 0 65536  1024  1024  256 0 1 10 0 $ch
run  2 0 0 1 65536  1024  1024  256 0 1 10 0 $ch
run  2 0 1 1  1024  1024 65536  256 1 0 10 0 $ch
done
# the modes it replaces, same shapes
run  0 2 0 0 65536  1024  1024  256 0 1 10
run  0 1 0 0 65536  1024  1024  256 0 1 10
run  1 1 0 0 65536  1024  1024  256 0 1 10
# wider K (config-4 size in fp32 parity mode)
run  2 0 0 0 16384  4096  4096  256 0 1 5
run  2 0 1 1  4096  4096 16384  256 1 0 5
grep -E "^==|RESULT|FAIL|exit=[1-9]|rel_fro|f16 pair" $LOG | tail -150
