#!/bin/bash
# full GPU suite + default bench line
D=gpurun_out/${1:-r2a}
mkdir -p $D
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $D/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $D/bench.json 2> $D/bench.err
tail -c 3000 $D/bench.json
tail -5 $D/bench.err
