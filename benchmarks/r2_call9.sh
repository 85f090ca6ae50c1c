set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pool.py -q -x -k "pool3d" > $O/r2_pool3d_tests.log 2>&1; tail -25 $O/r2_pool3d_tests.log
(cd benchmarks && timeout 300 python misc_paths.py pool3d 2>&1 | cut -c1-300) | tee $O/r2_pool3d_bench.txt
(cd benchmarks && KX_POOL3_NZC=1 timeout 300 python misc_paths.py pool3d 2>&1 | cut -c1-300) | tee -a $O/r2_pool3d_bench.txt
(cd benchmarks && KX_POOL3_NZC=4 timeout 300 python misc_paths.py pool3d 2>&1 | cut -c1-300) | tee -a $O/r2_pool3d_bench.txt
