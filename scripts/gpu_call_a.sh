#!/bin/bash
# full GPU suite + default bench line
D=gpurun_out/${1:-r2a}
mkdir -p $D
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee $D/smoke.txt
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
