set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2_gpu_tests_full.log 2>&1; tail -4 $O/r2_gpu_tests_full.log
timeout 300 python benchmarks/ref_harness.py > $O/r2_ref_harness.jsonl 2> $O/r2_ref_harness.err; tail -2 $O/r2_ref_harness.err; cut -c1-220 $O/r2_ref_harness.jsonl
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity_check','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for c in d['configs']: print(c['config'][:70], round(c['glups'],1), round(c['roofline_frac'],3), c['parity'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_reference_arm.json 2>/dev/null; cut -c1-300 $O/r2c_bench_reference_arm.json
