mkdir -p gpurun_out
for c in 4 6 8; do
TOPS_F16X3_FWD_HEAD=$c timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side --no-extra > gpurun_out/bq.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/bq.json'))
p=d['parity']
print("head $c", round(d['ms_per_step'],4), round(d['roofline']['per_kernel_ms']['gemm_fwd'],4), 'A %.2e dX %.2e dW %.2e db %.2e' % (p['A'],p['dX'],p['dW'],p['db']))
PY
done
