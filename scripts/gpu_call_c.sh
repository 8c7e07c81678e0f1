#!/bin/bash
mkdir -p gpurun_out/r2c
(timeout 300 python scripts/ozaki_prof.py 16 > gpurun_out/r2c/ozaki_prof.txt 2>&1; echo "exit $?" >> gpurun_out/r2c/ozaki_prof.txt)
cat gpurun_out/r2c/ozaki_prof.txt
(timeout 200 python scripts/ozaki_check.py 12 > gpurun_out/r2c/ozaki_check.txt 2>&1; echo "exit $?" >> gpurun_out/r2c/ozaki_check.txt)
tail -8 gpurun_out/r2c/ozaki_check.txt
