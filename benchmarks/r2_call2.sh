set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2_gpu_tests.log 2>&1
tail -60 gpurun_out/r2_gpu_tests.log
