mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bq.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bq.json'))
print(d['ms_per_step'], d['roofline']['per_kernel_ms'], d['parity']['ok'])
print(json.dumps(d['throughput_mode'])[:900])
PY
