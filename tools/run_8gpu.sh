mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --no-side > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read())
print(d['ms_per_step'], d['value'], d['config']['allreduce'][:80], d['config']['allreduce_trial'])
print(d['parity']['ok'], d['e2e']['value'], d['e2e']['cpu_affinity']); print(d['extra']['config4'])
PY
