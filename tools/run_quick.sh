for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-side --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'], d['parity']['ok'], d['clocks'])"
done
