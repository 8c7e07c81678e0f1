#!/bin/bash
# a library variant of the sliced contraction: correctness first, then timing against the default
D=gpurun_out/${1:-r2p16}
V=${2:-PROD16}
mkdir -p $D
L=$PWD/scripts/variants/libfqe_$V.so
(FQEB_B200_LIB=$L timeout -s KILL 200 python scripts/ozaki_check.py 16 > $D/check_$V.txt 2>&1; echo "exit $?" >> $D/check_$V.txt)
tail -5 $D/check_$V.txt
for v in default $V; do
  echo "=== $v" >> $D/variants.txt
  LL=$PWD/scripts/variants/libfqe_$v.so
  [ $v = default ] && LL=$PWD/openfermion-fqe_b200/fqe_b200/lib/libfqe_b200.so
  (FQEB_OZAKI_PROF=0 FQEB_B200_LIB=$LL timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/variants.txt 2>&1; echo "exit $?" >> $D/variants.txt)
done
(FQEB_OZAKI_PROF=1 FQEB_B200_LIB=$L timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/variants.txt 2>&1)
cat $D/variants.txt
