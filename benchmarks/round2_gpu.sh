#!/bin/bash
# The GPU commands behind the round-2 numbers in profiles/ (each block is one `gpurun` call; everything runs under its
# own `timeout`, results land in gpurun_out/).
#
#   gpurun --timeout 1500 -- 'bash benchmarks/round2_gpu.sh single'      # 1 GPU: tests, bench (both arms), harnesses
#   gpurun --timeout 900  -- 'bash benchmarks/round2_gpu.sh profile'     # 1 GPU: ncu launch list + full captures
#   gpurun --gpus N --timeout 900 -- "NG=N bash benchmarks/round2_gpu.sh multi"   # N = 2, 4, 8
set -u
mkdir -p gpurun_out
O=gpurun_out
case "${1:-single}" in
single)
  timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2_gpu_tests_full.log 2>&1; tail -3 $O/r2_gpu_tests_full.log
  timeout 600 python bench.py --steps 100 --warmup 5 > $O/r2d_bench_n1.json 2> $O/r2d_bench_n1.err; echo "bench rc=$?"
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2d_bench_reference_arm.json 2>/dev/null
  timeout 300 python benchmarks/ref_harness.py > $O/r2_ref_harness.jsonl 2>/dev/null
  (cd benchmarks && timeout 600 python misc_paths.py 2>&1 | cut -c1-400) > $O/r2_misc_paths.jsonl
  (cd benchmarks && timeout 300 python run_configs.py --reps 8 2>&1 | cut -c1-400) > $O/r2_configs_single_gpu.jsonl
  ;;
profile)
  # launch list of the bench command itself (run plain first, exit 0), then one full capture of the dominant kernel
  timeout 300 python bench.py --steps 20 --warmup 3 --no-configs > $O/r2d_bench_plain.json 2>/dev/null || exit 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv \
      python bench.py --steps 20 --warmup 3 --no-configs > /dev/null 2>&1
  NCU="ncu --set full --clock-control none --import-source on"
  timeout 300 $NCU -k regex:stencil2d_kernel -s 8 -c 1 -o $O/p2_stencil2d python benchmarks/run_configs.py --only C2 --reps 2 > /dev/null 2>&1
  timeout 300 $NCU -k regex:stencil3d_star -s 3 -c 1 -o $O/p2_stencil3d_star python benchmarks/run_configs.py --only C5 --reps 2 > /dev/null 2>&1
  (cd benchmarks && timeout 300 $NCU -k regex:pool3d -s 4 -c 1 -o ../$O/p2_pool3d python misc_paths.py pool3d > /dev/null 2>&1)
  (cd benchmarks && timeout 300 $NCU -k regex:stencil2d_mesh -s 4 -c 1 -o ../$O/p2_mesh python misc_paths.py mesh > /dev/null 2>&1)
  ls -la $O/p2_*.ncu-rep
  ;;
multi)
  NG=${NG:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
  nvidia-smi topo -m > $O/r2_topo_n$NG.txt 2>&1; nproc >> $O/r2_topo_n$NG.txt; lscpu | grep -i "numa\|socket\|model name" >> $O/r2_topo_n$NG.txt
  timeout 400 $TR --master-port 29511 tests/dist_gpu_check.py > $O/r2_dist_check_n$NG.log 2>&1; echo "dist_gpu_check rc=$?"
  for tr in p2p nccl; do
    timeout 200 $TR --master-port 29514 benchmarks/step_timeline.py --transport $tr --steps 200 2> /dev/null | tail -1 > $O/r2_timeline_n${NG}_$tr.json
  done
  timeout 500 $TR --master-port 29512 bench.py --gpus $NG --steps 300 --warmup 5 > $O/r2d_bench_n$NG.json 2> $O/r2d_bench_n$NG.err; echo "bench rc=$?"
  timeout 200 $TR --master-port 29513 bench.py --impl reference --gpus $NG --steps 2 --warmup 1 > $O/r2d_bench_reference_arm_n$NG.json 2>/dev/null
  ;;
esac
