mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma_kernel -c 6 -f -o gpurun_out/prof_cfg3_gemm python tools/bench_configs.py 3 --prec f16x3 > gpurun_out/ncu_cfg3.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_cfg3_gemm.ncu-rep gpurun_out/r2_cfg3_gemm_ncu > /dev/null 2>&1
ls -la gpurun_out/prof_cfg3_gemm.ncu-rep
