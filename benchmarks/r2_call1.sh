set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > $O/r2_box.txt; nproc >> $O/r2_box.txt
export KX_TEST_EXPERIMENTAL=1
timeout 150 python -m pytest tests/test_gpu_fastpaths.py -q -k "pair or plane_stationary or folded_1d" > $O/r2_experimental_tests.log 2>&1
tail -5 $O/r2_experimental_tests.log
(
  cd benchmarks
  echo "== C5 star kernel: default vs KX_STAR_PAIR=1"
  timeout 120 python run_configs.py --only C5 --reps 6 2>/dev/null | cut -c1-300
  KX_STAR_PAIR=1 timeout 120 python run_configs.py --only C5 --reps 6 2>/dev/null | cut -c1-300
  echo "== dense 27-point"
  timeout 120 python misc_paths.py dense3d 2>&1 | cut -c1-300
  echo "== 1-D FIR: one-row launch vs KX_FOLD_1D=1"
  timeout 60 python misc_paths.py fir 2>&1 | cut -c1-300
  KX_FOLD_1D=1 timeout 60 python misc_paths.py fir 2>&1 | cut -c1-300
) > $O/r2_variants.txt 2>&1
cat $O/r2_variants.txt
unset KX_TEST_EXPERIMENTAL
timeout 300 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_fastpaths.py -q -x -k 'vjp or gradient or conv_rows' > $O/r2_vjp_tests.log 2>&1; tail -15 $O/r2_vjp_tests.log
timeout 200 python bench.py --steps 100 --warmup 5 > $O/r2a_bench_n1.json 2> $O/r2a_bench_n1.err; tail -c 1500 $O/r2a_bench_n1.json
NCU="ncu --set full --clock-control none --import-source on"
(cd benchmarks && timeout 180 $NCU -k regex:stencil2d_mesh -s 4 -c 1 -o ../$O/p_mesh2 python misc_paths.py mesh > /dev/null 2>&1)
ls -la $O/*.ncu-rep 2>/dev/null | tail -5
