set -u
mkdir -p gpurun_out
O=gpurun_out
cd benchmarks
for pf in 0 2 4 8; do echo "== PF=$pf"; KX_S2_PF=$pf timeout 200 python misc_paths.py wide 2>&1 | grep -i "gauss\|dense 7x7\|moving" | cut -c1-200; done | tee ../$O/r2_wide_pf.txt
