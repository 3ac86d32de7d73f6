mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout 300 python tools/bench_configs.py 3 5 --prec f16x3 2>&1 | tee gpurun_out/cfg35.log | cut -c1-700
timeout 600 python bench.py > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; cut -c1-1500 gpurun_out/bench_q.json
