#!/bin/bash
# N-GPU bench: the default line (det sharding, overlapped all-reduce, pair leg, secondary leg) and
# the same det run with one all-reduce after the build
D=gpurun_out/${1:-r2m8}
N=${2:-8}
mkdir -p $D
run() {  # name, env, args
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 $3 > $D/$1.json 2> $D/$1.err
  python - <<PY
import json
for l in open("$D/$1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$1", "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3),
              "verify", d.get("verify_rel_err"), "shard_verify", d.get("shard_verify_rel_err"),
              "pair", (d.get("pair_shard") or {}).get("value"), "secondary", (d.get("secondary") or {}).get("value"))
PY
  tail -2 $D/$1.err
}
run default_n$N FQEB_ALLREDUCE_SLICES=4 "--no-cpu-baseline"
run det_slices1_n$N FQEB_ALLREDUCE_SLICES=1 "--no-secondary --no-cpu-baseline"
