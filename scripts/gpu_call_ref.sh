#!/bin/bash
# reference arm on the GPU box's host cores + the tests touched last
D=gpurun_out/${1:-r2ref}
mkdir -p $D
python -m pytest tests/test_gpu_kernels.py tests/test_cabi_symbols.py -q -m gpu 2>&1 | tail -4 | tee $D/pytest_part.txt
timeout -s KILL 700 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
tail -c 1800 $D/bench_reference.json; tail -3 $D/bench_reference.err
