set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/dist_gpu_check.py > $O/r2_dist_check_n2.log 2>&1; echo "dist_gpu_check rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*" $O/r2_dist_check_n2.log | tail -30
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > $O/r2_bench_n2_p2p.json 2> $O/r2_bench_n2_p2p.err; echo "bench p2p rc=$?"; cut -c1-900 $O/r2_bench_n2_p2p.json
KX_DIST_TRANSPORT=nccl timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 > $O/r2_bench_n2_nccl.json 2> $O/r2_bench_n2_nccl.err; echo "bench nccl rc=$?"; cut -c1-900 $O/r2_bench_n2_nccl.json
tail -5 $O/r2_bench_n2_p2p.err
