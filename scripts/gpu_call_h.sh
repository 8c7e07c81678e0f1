#!/bin/bash
# A/B of producer variants of the sliced contraction + MMA issue microbenchmark + RDM tests
D=gpurun_out/${1:-r2w}
mkdir -p $D
(timeout -s KILL 120 python scripts/ozaki_check.py 8 > $D/ozaki_check8.txt 2>&1; echo "exit $?" >> $D/ozaki_check8.txt)
tail -4 $D/ozaki_check8.txt
grep -q "^exit 0" $D/ozaki_check8.txt || exit 1
(timeout -s KILL 300 python scripts/ozaki_check.py 16 > $D/ozaki_check.txt 2>&1; echo "exit $?" >> $D/ozaki_check.txt)
tail -9 $D/ozaki_check.txt
for v in default $VARIANTS; do
  echo "=== $v" >> $D/variants.txt
  L=$PWD/scripts/variants/libfqe_$v.so
  [ $v = default ] && L=$PWD/openfermion-fqe_b200/fqe_b200/lib/libfqe_b200.so
  (FQEB_OZAKI_PROF=0 FQEB_B200_LIB=$L timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/variants.txt 2>&1; echo "exit $?" >> $D/variants.txt)
done
cat $D/variants.txt
(timeout -s KILL 200 python scripts/ozaki_prof.py 16 > $D/ozaki_prof.txt 2>&1; echo "exit $?" >> $D/ozaki_prof.txt)
cat $D/ozaki_prof.txt
(cd profiles/microbench && timeout -s KILL 120 ./i8_tmem_a > ../../$D/i8_tmem_a.txt 2>&1; echo "exit $?" >> ../../$D/i8_tmem_a.txt)
grep -i "rate\|exit" $D/i8_tmem_a.txt
python -m pytest tests/test_gpu_rdm.py -x -q 2>&1 | tail -5 | tee $D/pytest_rdm.txt
