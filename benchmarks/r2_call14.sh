set -u
mkdir -p gpurun_out
O=gpurun_out
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29512 bench.py --gpus $NG --steps 200 --warmup 5 > $O/r2c_bench_n$NG.json 2> $O/r2c_bench_n$NG.err; echo "bench rc=$?"
tail -3 $O/r2c_bench_n$NG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_n$NG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity_check','clocks')})
print(d['config']['parallelism'])
print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for c in d['configs']: print(json.dumps(c)[:500])
PY
timeout 200 $TR --master-port 29513 bench.py --impl reference --gpus $NG --steps 2 --warmup 1 | cut -c1-200
