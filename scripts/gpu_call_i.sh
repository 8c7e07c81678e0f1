#!/bin/bash
# MMA issue microbenchmark (two issuing warps), scatter occupancy experiment, bench line
D=gpurun_out/${1:-r2x}
mkdir -p $D
(cd profiles/microbench && timeout -s KILL 120 ./i8_tmem_a > ../../$D/i8_tmem_a.txt 2>&1; echo "exit $?" >> ../../$D/i8_tmem_a.txt)
grep -i "rate\|exit\|drain" $D/i8_tmem_a.txt
for k in 0 1 2 4; do
  echo "=== FQEB_SCATTER_CTAS_PER_SM=$k" >> $D/scatter_occ.txt
  (FQEB_OZAKI_PROF=0 FQEB_SCATTER_CTAS_PER_SM=$k timeout -s KILL 200 python scripts/ozaki_prof.py 16 >> $D/scatter_occ.txt 2>&1; echo "exit $?" >> $D/scatter_occ.txt)
done
cat $D/scatter_occ.txt
python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $D/bench.json 2> $D/bench.err
python - <<PY
import json
for l in open("$D/bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", {k: v for k, v in d["e2e"].items() if "value" in k})
PY
tail -3 $D/bench.err
