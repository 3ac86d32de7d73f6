mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "fflayer or config2 or cta_pair or host" > gpurun_out/pytest_q.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_q.log
tail -3 gpurun_out/pytest_q.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side --no-extra > gpurun_out/bq.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print(d['ms_per_step'], d['roofline']['per_kernel_ms'], sum(d['roofline']['per_kernel_ms'].values()), d['parity']['ok'], d['gpu_launches'])
PY
