#!/bin/bash
mkdir -p gpurun_out/r2d
FQEB_OZAKI_PROF=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sigma_ozaki -s 3 -c 1 -o gpurun_out/r2d/ozaki -f python scripts/ozaki_prof.py 16 > gpurun_out/r2d/ncu_log.txt 2>&1
tail -5 gpurun_out/r2d/ncu_log.txt
ls -la gpurun_out/r2d/
