#!/bin/bash
# ncu --set full of the three GEMMs of one bench step (after the warm-up launches) + the launch list of the same command
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 9 -c 3 -f -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-side > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-side > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
