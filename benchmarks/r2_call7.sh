set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:stencil3d_star -s 3 -c 1 -o gpurun_out/p2_star_groups python benchmarks/run_configs.py --only C5 --reps 2 > /dev/null 2>&1
KX_STAR_GRID3D=1 $NCU -k regex:stencil3d_star -s 3 -c 1 -o gpurun_out/p2_star_grid3d python benchmarks/run_configs.py --only C5 --reps 2 > /dev/null 2>&1
ls -la gpurun_out/p2_*.ncu-rep
