#!/bin/bash
# sliced contraction: all-sector check against the FP64 paths, then phase times + cycle counters
D=gpurun_out/${1:-r2e}
mkdir -p $D
(timeout 400 python scripts/ozaki_check.py 16 > $D/ozaki_check.txt 2>&1; echo "exit $?" >> $D/ozaki_check.txt)
tail -34 $D/ozaki_check.txt
(timeout 300 python scripts/ozaki_prof.py 16 > $D/ozaki_prof.txt 2>&1; echo "exit $?" >> $D/ozaki_prof.txt)
cat $D/ozaki_prof.txt
