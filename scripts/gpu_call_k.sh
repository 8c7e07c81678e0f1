#!/bin/bash
# multi-GPU: det sharding with / without the overlapped all-reduce, pair sharding; N = $2 GPUs
D=gpurun_out/${1:-r2z}
N=${2:-2}
mkdir -p $D
python -m pytest tests/test_gpu_sigma.py -x -q -m gpu -k deferred 2>&1 | tail -3 | tee $D/pytest_deferred.txt
run() {  # name, extra env, extra args
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-cpu-baseline \
    --verify $3 > $D/$1.json 2> $D/$1.err
  python - <<PY
import json
for l in open("$D/$1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$1", "value", round(d["value"], 3), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3),
              "verify", d.get("verify_rel_err"), "shard_verify", d.get("shard_verify_rel_err"),
              "pair", (d.get("pair_shard") or {}).get("value"))
PY
  tail -2 $D/$1.err
}
run det_slices4 FQEB_ALLREDUCE_SLICES=4 "--shard det"
run det_slices1 FQEB_ALLREDUCE_SLICES=1 "--shard det"
