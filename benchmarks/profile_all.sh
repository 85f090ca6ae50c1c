#!/bin/bash
# ncu captures of every roofline-tuned kernel on its BASELINE config (run under gpurun, 1 GPU).
# Each kernel is profiled only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python benchmarks/run_configs.py --reps 6 > gpurun_out/configs_plain.jsonl 2> gpurun_out/configs_plain.err || exit 1
$NCU -k regex:stencil2d -s 8 -c 1 -o gpurun_out/p_stencil2d python benchmarks/run_configs.py --only C2 --reps 2 > /dev/null 2>&1
$NCU -k regex:stencil3d_star -s 3 -c 1 -o gpurun_out/p_stencil3d_star python benchmarks/run_configs.py --only C5 --reps 2 > /dev/null 2>&1
$NCU -k regex:scan_rowmarch -s 70 -c 1 -o gpurun_out/p_rowmarch python benchmarks/run_configs.py --only C4 --reps 2 > /dev/null 2>&1
$NCU -k regex:conv_fwd_rows -s 8 -c 1 -o gpurun_out/p_conv_fwd python benchmarks/run_configs.py --only C3a --reps 2 > /dev/null 2>&1
$NCU -k regex:conv_wgrad_rows -s 3 -c 1 -o gpurun_out/p_conv_wgrad python benchmarks/run_configs.py --only C3a --reps 2 > /dev/null 2>&1
$NCU -k regex:patch2d -s 8 -c 1 -o gpurun_out/p_patch2d python benchmarks/run_configs.py --only C3b --reps 2 > /dev/null 2>&1
(cd benchmarks && python misc_paths.py > ../gpurun_out/misc_paths.jsonl 2> ../gpurun_out/misc_paths.err)
(cd benchmarks && $NCU -k regex:pool2d -s 12 -c 1 -o ../gpurun_out/p_pool2d_s2 python misc_paths.py pool > /dev/null 2>&1)
# stride-1 window classes added in r1c (separable folds, large windows), dense 3-D on the conv ring, the mesh kernel, 1-D FIR
(cd benchmarks && python misc_paths.py wide,dense3d,mesh,fir > ../gpurun_out/misc_paths_r1c.jsonl 2> ../gpurun_out/misc_paths_r1c.err)
(cd benchmarks && $NCU -k regex:stencil2d_kernel -s 20 -c 1 -o ../gpurun_out/p_s2_gauss7 python misc_paths.py wide > /dev/null 2>&1)   # first timed launch of the 7x7 separable case
(cd benchmarks && $NCU -k regex:stencil2d_mesh -s 4 -c 1 -o ../gpurun_out/p_mesh python misc_paths.py mesh > /dev/null 2>&1)
(cd benchmarks && $NCU -k regex:conv_fwd_rows -s 4 -c 1 -o ../gpurun_out/p_dense3d python misc_paths.py dense3d > /dev/null 2>&1)
ls -la gpurun_out/p_*.ncu-rep
