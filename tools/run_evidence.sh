#!/bin/bash
# One GPU visit producing the artefacts summarised under profiles/: ncu captures, launch list, bench lines, parity report,
# bandwidth table, sanitizer logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
bash tools/run_ncu_gemm.sh > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep gpurun_out/r2_gemm_f16x3_ncu > /dev/null 2>&1
python tools/ncu_hot_sass.py gpurun_out/prof_gemm.ncu-rep 0 40 > gpurun_out/r2_gemm_fwd_hot_sass.txt 2>&1
rm -f gpurun_out/prof_gemm.ncu-rep          # gpurun brings back at most 64 MiB: keep the summaries, not the reports
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 300 python tools/parity_report.py > gpurun_out/parity.log 2>&1
timeout 600 python tools/bench_bandwidth.py > gpurun_out/bench_bandwidth.log 2>&1
timeout 600 python tools/bench_configs.py 3 5 > gpurun_out/bench_configs_3_5.log 2>&1
for k in "gemv 1024x1024:k_gemv_rows:4" "split_f16_rows:k_split_f16_rows:1" "ger 16384:k_outer_rows:4" "sum_rows:k_colsum_partial:4" "transp rank 3:k_permute_tiled:4" "logistic'(x):k_map2:4"; do
  name="${k%%:*}"; rest="${k#*:}"; kern="${rest%%:*}"; skip="${rest##*:}"
  timeout 600 ncu --set full --clock-control none -k regex:$kern -s $skip -c 1 -f -o gpurun_out/prof_bw_$kern python tools/bench_bandwidth.py --one "$name" > gpurun_out/ncu_bw_$kern.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_bw_$kern.ncu-rep gpurun_out/r2_bw_${kern}_ncu > /dev/null 2>&1
  rm -f gpurun_out/prof_bw_$kern.ncu-rep
done
# config 3: the six tcgen05 GEMMs of one netGrad step and the three skinny output-layer kernels, launch list of the whole step
timeout 900 ncu --set full --clock-control none -k regex:gemm_umma_kernel -c 6 -f -o gpurun_out/prof_cfg3_gemm python tools/bench_configs.py 3 --prec f16x3 > gpurun_out/ncu_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_cfg3_gemm.ncu-rep gpurun_out/r2_cfg3_gemm_ncu > /dev/null 2>&1
rm -f gpurun_out/prof_cfg3_gemm.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:k_skinny -c 3 -f -o gpurun_out/prof_skinny python tools/bench_configs.py 3 --prec f16x3 > gpurun_out/ncu_skinny.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_skinny.ncu-rep gpurun_out/r2_skinny_ncu > /dev/null 2>&1
rm -f gpurun_out/prof_skinny.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_config3_f16x3.csv python tools/bench_configs.py 3 --prec f16x3 > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1
tail -n 3 gpurun_out/sanitizer_memcheck.log; tail -n 3 gpurun_out/sanitizer_racecheck.log; ls gpurun_out; du -sh gpurun_out
