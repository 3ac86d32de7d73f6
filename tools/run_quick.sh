mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in "2 4" "2 0" "2 6"; do set -- $cfg
echo "== fwd chunk $1 head $2"
TOPS_F16X3_FWD_CHUNK=$1 TOPS_F16X3_FWD_HEAD=$2 timeout 300 python tools/parity_report.py 2>&1 | grep f16x3
TOPS_F16X3_FWD_CHUNK=$1 TOPS_F16X3_FWD_HEAD=$2 timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done
