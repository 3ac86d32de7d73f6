mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_q.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_q.log
tail -4 gpurun_out/pytest_q.log
timeout 300 python tools/bench_configs.py 3 --prec f16x3 2>&1 | tee gpurun_out/cfg3.log | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3.csv python tools/bench_configs.py 3 --prec f16x3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg3.csv')) if len(r)>5 and r[0].isdigit()]
print(len(rows))
names=collections.Counter(r[4][:60] for r in rows)
for k,v in names.most_common(30): print(v,k)
PY
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'], d['parity'])"
