#!/bin/bash
# sliced contraction: small-sector check first (short timeout: a hang must not eat the budget),
# then all sectors against the FP64 paths, then phase times + cycle counters
D=gpurun_out/${1:-r2e}
mkdir -p $D
(timeout -s KILL 120 python scripts/ozaki_check.py 8 > $D/ozaki_check8.txt 2>&1; echo "exit $?" >> $D/ozaki_check8.txt)
tail -24 $D/ozaki_check8.txt
grep -q "^exit 0" $D/ozaki_check8.txt || exit 1
(timeout -s KILL 300 python scripts/ozaki_check.py 16 > $D/ozaki_check.txt 2>&1; echo "exit $?" >> $D/ozaki_check.txt)
tail -9 $D/ozaki_check.txt
(timeout -s KILL 200 python scripts/ozaki_prof.py 16 > $D/ozaki_prof.txt 2>&1; echo "exit $?" >> $D/ozaki_prof.txt)
cat $D/ozaki_prof.txt
