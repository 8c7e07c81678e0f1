#!/bin/bash
mkdir -p gpurun_out/r2b
(timeout 300 python scripts/ozaki_check.py 16 > gpurun_out/r2b/ozaki_check.txt 2>&1; echo "exit $?" >> gpurun_out/r2b/ozaki_check.txt)
tail -40 gpurun_out/r2b/ozaki_check.txt
if ! grep -q "^exit 0" gpurun_out/r2b/ozaki_check.txt; then
  (timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/ozaki_check.py 4 > gpurun_out/r2b/sanitizer.txt 2>&1; echo "exit $?" >> gpurun_out/r2b/sanitizer.txt)
  grep -v "^=========     at\|^=========         at" gpurun_out/r2b/sanitizer.txt | head -60
fi
