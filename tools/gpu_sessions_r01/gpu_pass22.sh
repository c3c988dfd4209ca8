#!/bin/bash
mkdir -p gpurun_out
# full capture of one mid-factorization launch of each leaf kernel (skip the first launches: small trailing sizes come last, large first)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tri_leaf_kernel' -s 40 -c 1 -o gpurun_out/p22_tri_leaf -f python tools/prof_lapack.py potrf 8192 > gpurun_out/p22_a.log 2>&1; tail -1 gpurun_out/p22_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'potf2_leaf_kernel' -s 10 -c 1 -o gpurun_out/p22_potf2 -f python tools/prof_lapack.py potrf 8192 > gpurun_out/p22_b.log 2>&1; tail -1 gpurun_out/p22_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'getf2_panel_kernel' -s 10 -c 1 -o gpurun_out/p22_getf2 -f python tools/prof_lapack.py getrf 8192 > gpurun_out/p22_c.log 2>&1; tail -1 gpurun_out/p22_c.log
ls -la gpurun_out/*.ncu-rep
