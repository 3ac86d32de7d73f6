mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fflayer or cta_pair or host or config2" > gpurun_out/pytest_ff.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ff.log
tail -3 gpurun_out/pytest_ff.log
for d in 0 1; do
TOPS_GEMM_DEBUG=$d timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'], d['parity']['ok'])"
done
