mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/run_ncu_gemm.sh > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep gpurun_out/r2_gemm_f16x3_ncu > /dev/null 2>&1
python tools/ncu_hot_sass.py gpurun_out/prof_gemm.ncu-rep 0 40 > gpurun_out/r2_gemm_fwd_hot_sass.txt 2>&1
rm -f gpurun_out/prof_gemm.ncu-rep
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read()); print(d['ms_per_step'], d['value'], d['roofline']['per_kernel_ms'], d['roofline']['frac'], d['parity']['ok'])"
grep -E "duration|tensor pipe active % \(elapsed" gpurun_out/r2_gemm_f16x3_ncu.md | cut -c1-160
