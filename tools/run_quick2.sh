mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_skinny -c 3 -f -o gpurun_out/prof_skinny python tools/bench_configs.py 3 --prec f16x3 > gpurun_out/ncu_skinny.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_skinny.ncu-rep gpurun_out/r2_skinny_ncu > /dev/null 2>&1
ncu -i gpurun_out/prof_skinny.ncu-rep --page details --csv > gpurun_out/skinny_details.csv 2>/dev/null
ls -la gpurun_out/prof_skinny.ncu-rep
