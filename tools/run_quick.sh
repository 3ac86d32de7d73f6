mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "fflayer or cta_pair or host" > gpurun_out/pytest_ff.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ff.log
tail -15 gpurun_out/pytest_ff.log
timeout 300 python tools/parity_report.py > gpurun_out/parity.log 2>&1; cat gpurun_out/parity.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
