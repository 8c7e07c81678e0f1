#!/bin/bash
# end-of-round validation: smoke, full GPU suite, the default bench line, the reference arm
D=gpurun_out/${1:-r2final}
mkdir -p $D
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $D/smoke.txt
python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $D/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $D/bench.json 2> $D/bench.err
tail -c 600 $D/bench.json; tail -3 $D/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
tail -c 1500 $D/bench_reference.json; tail -3 $D/bench_reference.err
