set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastpaths.py -q -x -k "dense3d" 2>&1 | tail -3
cd benchmarks
( timeout 300 python misc_paths.py dense3d 2>&1 | grep "dense 27" | cut -c1-260
  echo "== no groups"; KX_DENSE3D_NO_GROUPS=1 timeout 300 python misc_paths.py dense3d 2>&1 | grep "dense 27" | cut -c1-260
  echo "== zc 32"; KX_DENSE3D_ZC=32 timeout 300 python misc_paths.py dense3d 2>&1 | grep "dense 27" | cut -c1-260
  echo "== zc 128"; KX_DENSE3D_ZC=128 timeout 300 python misc_paths.py dense3d 2>&1 | grep "dense 27" | cut -c1-260
) | tee ../$O/r2_dense3d_bench.txt
