#!/bin/bash
# First GPU call of the next round (1 GPU): try the opt-in variants written at the end of round 1
# and capture the kernels the round-1 numbers point at.  Every command runs under its own `timeout`; results land
# in gpurun_out/.   gpurun --timeout 900 -- 'bash benchmarks/round2_first.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
export KX_TEST_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_gpu_fastpaths.py -q -k "pair or plane_stationary or folded_1d" > $O/r2_experimental_tests.log 2>&1
tail -3 $O/r2_experimental_tests.log
(
  cd benchmarks
  echo "== C5 star kernel: default vs KX_STAR_PAIR=1"
  timeout 120 python run_configs.py --only C5 --reps 6 2>/dev/null | cut -c1-200
  KX_STAR_PAIR=1 timeout 120 python run_configs.py --only C5 --reps 6 2>/dev/null | cut -c1-200
  echo "== dense 27-point: conv row ring vs plane-stationary kernel (printed by misc_paths under KX_TEST_EXPERIMENTAL)"
  timeout 120 python misc_paths.py dense3d 2>&1 | cut -c1-220
  echo "== 1-D FIR: one-row launch vs KX_FOLD_1D=1"
  timeout 60 python misc_paths.py fir 2>&1 | cut -c1-200
  KX_FOLD_1D=1 timeout 60 python misc_paths.py fir 2>&1 | cut -c1-200
) > $O/r2_variants.txt 2>&1
cat $O/r2_variants.txt
NCU="ncu --set full --clock-control none --import-source on"
(cd benchmarks && timeout 180 $NCU -k regex:stencil2d_kernel -s 20 -c 1 -o ../$O/p_s2_gauss7 python misc_paths.py wide > /dev/null 2>&1)
(cd benchmarks && timeout 180 $NCU -k regex:stencil2d_mesh -s 4 -c 1 -o ../$O/p_mesh2 python misc_paths.py mesh > /dev/null 2>&1)
(cd benchmarks && timeout 180 $NCU -k regex:conv_fwd_rows -s 4 -c 1 -o ../$O/p_dense3d python misc_paths.py dense3d > /dev/null 2>&1)
ls -la $O/p_*.ncu-rep 2>/dev/null | tail -5
# 2 GPUs (separate call, charged twice):
#   gpurun --gpus 2 --timeout 300 -- 'KX_TEST_EXPERIMENTAL=1 timeout 200 python -m torch.distributed.run --nnodes=1 \
#       --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py'
