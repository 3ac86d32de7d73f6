mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_recurrent.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_graph.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_graph.log
tail -25 gpurun_out/pytest_graph.log
