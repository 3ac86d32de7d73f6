mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_par.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_par.log
tail -5 gpurun_out/pytest_par.log
timeout 900 python tools/bench_bandwidth.py > gpurun_out/bench_bandwidth.log 2>&1; cat gpurun_out/bench_bandwidth.log | tail -25
