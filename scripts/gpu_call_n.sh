#!/bin/bash
# N-GPU det-sharded bench, overlapped all-reduce on a high-priority communicator
D=gpurun_out/${1:-r2o8}
N=${2:-8}
mkdir -p $D
for sl in ${SLICES:-6}; do
  env FQEB_ALLREDUCE_SLICES=$sl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-cpu-baseline \
    > $D/det_slices$sl.json 2> $D/det_slices$sl.err
  python - <<PY
import json
for l in open("$D/det_slices$sl.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("slices $sl", "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3),
              "verify", d.get("verify_rel_err"), {k: round(v["ms_per_step"], 2) for k, v in d["phases"].items()})
PY
  tail -2 $D/det_slices$sl.err
done
