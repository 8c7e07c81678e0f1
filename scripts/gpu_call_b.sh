#!/bin/bash
mkdir -p gpurun_out/r2b
(timeout 120 profiles/microbench/i8_umma > gpurun_out/r2b/i8_umma.txt 2>&1; echo "exit $?" >> gpurun_out/r2b/i8_umma.txt)
cat gpurun_out/r2b/i8_umma.txt


(timeout 300 python scripts/ozaki_check.py 16 > gpurun_out/r2b/ozaki_check.txt 2>&1; echo "exit $?" >> gpurun_out/r2b/ozaki_check.txt)
tail -40 gpurun_out/r2b/ozaki_check.txt
