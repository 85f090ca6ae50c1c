set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; echo "bench rc=$?"
tail -3 $O/r2b_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity_check','clocks')})
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for c in d['configs']: print(json.dumps(c)[:400])
PY
