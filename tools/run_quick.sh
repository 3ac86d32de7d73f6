mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_q.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_q.log
tail -4 gpurun_out/pytest_q.log
timeout 300 python tools/bench_configs.py 5 --prec f16x3 2>&1 | tail -1 | cut -c1-1100
