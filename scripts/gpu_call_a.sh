#!/bin/bash
# first GPU call of round 2: tcgen05 INT8 microbenchmark, large-golden parity tests, full GPU suite, bench
mkdir -p gpurun_out/r2a
(timeout 120 profiles/microbench/i8_umma > gpurun_out/r2a/i8_umma.txt 2>&1; echo "exit $?" >> gpurun_out/r2a/i8_umma.txt)
cat gpurun_out/r2a/i8_umma.txt
python -m pytest tests/test_gpu_large_goldens.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2a/large_goldens.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2a/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
tail -c 2500 gpurun_out/r2a/bench.json
