#!/bin/bash
# timing experiments: variant builds of the library with parts of the sliced kernel compiled out
D=gpurun_out/${1:-r2f}
mkdir -p $D
for v in NOSTORE NOLOAD NOALPHA NOBETA; do
  echo "=== $v" >> $D/variants.txt
  (FQEB_B200_LIB=$PWD/scripts/variants/libfqe_$v.so timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/variants.txt 2>&1; echo "exit $?" >> $D/variants.txt)
done
cat $D/variants.txt
