mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_graph.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_graph.log
tail -4 gpurun_out/pytest_graph.log
timeout 600 python tools/bench_configs.py 5 --prec f16x3 > gpurun_out/bench_configs_5.log 2>&1; cat gpurun_out/bench_configs_5.log
