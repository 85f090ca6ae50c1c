set -u
mkdir -p gpurun_out
O=gpurun_out
NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/r2_topo_n$NG.txt 2>&1; nproc >> $O/r2_topo_n$NG.txt; lscpu | grep -i "numa\|socket\|model name" >> $O/r2_topo_n$NG.txt
timeout 400 $TR --master-port 29511 tests/dist_gpu_check.py > $O/r2_dist_check_n$NG.log 2>&1; echo "dist_gpu_check rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*\|OMP_NUM" $O/r2_dist_check_n$NG.log | tail -34
for tr in p2p nccl; do
  timeout 200 $TR --master-port 29514 benchmarks/step_timeline.py --transport $tr --steps 200 2> $O/r2_timeline_n${NG}_$tr.err | tail -1 > $O/r2_timeline_n${NG}_$tr.json; echo "timeline $tr rc=$?"; cut -c1-1500 $O/r2_timeline_n${NG}_$tr.json
done
timeout 500 $TR --master-port 29512 bench.py --gpus $NG --steps 300 --warmup 5 > $O/r2b_bench_n$NG.json 2> $O/r2b_bench_n$NG.err; echo "bench rc=$?"
tail -3 $O/r2b_bench_n$NG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2b_bench_n$NG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','parity_check','clocks')})
print(d['config']['parallelism'])
print('roofline', d['roofline']['frac'], 'e2e', d['e2e'])
for c in d['configs']: print(json.dumps(c)[:700])
PY
