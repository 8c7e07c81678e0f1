#!/bin/bash
# end-of-round evidence: full GPU suite, default bench line, launch list of the bench
D=gpurun_out/${1:-r2final2}
mkdir -p $D
python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $D/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $D/bench.json 2> $D/bench.err
python - <<PY
import json
for l in open("$D/bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", {k: v for k, v in d["e2e"].items() if "value" in k},
              "frac", d["roofline"]["frac"], "verify", d["verify_rel_err"], {k: round(v["ms_per_step"], 2) for k, v in d["phases"].items()})
PY
tail -3 $D/bench.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $D/launches.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline \
  > $D/bench_under_ncu.log 2>&1
tail -2 $D/launches.csv | cut -c1-200
