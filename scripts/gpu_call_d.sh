#!/bin/bash
# ncu --set full capture of one launch of the sliced contraction kernel at norb=16
D=gpurun_out/${1:-r2d}
mkdir -p $D
FQEB_OZAKI_PROF=0 timeout -s KILL 700 ncu --set full --clock-control none --import-source on -k regex:k_sigma_ozaki -s 4 -c 1 -o $D/ozaki -f python scripts/ozaki_prof.py 16 > $D/ncu_log.txt 2>&1
tail -5 $D/ncu_log.txt
ls -la $D/
