#!/bin/bash
mkdir -p gpurun_out
for w in potrf getrf; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'potf2|getf2|perm_|tri_leaf|dmma|simt|window' --csv --log-file gpurun_out/p19_launches_$w.csv python tools/prof_lapack.py $w 8192 > gpurun_out/p19_$w.log 2>&1
  tail -2 gpurun_out/p19_$w.log
done
