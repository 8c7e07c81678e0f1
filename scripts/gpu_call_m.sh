#!/bin/bash
# A/B of library variants (scripts/variants/libfqe_<V>.so) of the sliced contraction
D=gpurun_out/${1:-r2n}
mkdir -p $D
for v in default $VARIANTS; do
  echo "=== $v" >> $D/variants.txt
  L=$PWD/scripts/variants/libfqe_$v.so
  [ $v = default ] && L=$PWD/openfermion-fqe_b200/fqe_b200/lib/libfqe_b200.so
  (FQEB_OZAKI_PROF=0 FQEB_B200_LIB=$L timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/variants.txt 2>&1; echo "exit $?" >> $D/variants.txt)
done
cat $D/variants.txt
