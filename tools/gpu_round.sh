#!/bin/bash
# One GPU visit: parity tests, bench (both precision modes), library peaks, ncu launch list + full capture of the GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 300 python tools/measure_peaks.py > gpurun_out/peaks.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 9 -c 3 -f -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
