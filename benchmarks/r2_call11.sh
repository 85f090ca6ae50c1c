set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_parity.py tests/test_gpu_interface.py -q -x 2>&1 | tail -5
cd benchmarks
( timeout 300 python misc_paths.py mesh 2>&1 | cut -c1-260
  echo "== nouniform"; KX_MESH_NO_UNIFORM=1 timeout 300 python misc_paths.py mesh 2>&1 | grep "kmap mesh" | cut -c1-260
  echo "== natural order"; KX_MESH_NATURAL_ORDER=1 timeout 300 python misc_paths.py mesh 2>&1 | grep "kmap mesh" | cut -c1-260; echo "== TH 8/16/24"; for th in 8 16 24; do KX_MESH_TH=$th timeout 300 python misc_paths.py mesh 2>&1 | grep "kmap mesh" | cut -c1-260; done
) | tee ../$O/r2_mesh_bench5.txt


