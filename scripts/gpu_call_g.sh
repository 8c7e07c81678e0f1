#!/bin/bash
# round-2 evidence pass: full GPU suite, default bench line, launch list of the bench, and one
# `ncu --set full` capture each of the sliced contraction kernel and of the scatter
D=gpurun_out/${1:-r2v}
mkdir -p $D
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $D/gpu_box.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $D/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $D/bench.json 2> $D/bench.err
tail -c 1500 $D/bench.json
tail -3 $D/bench.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $D/launches.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline \
  > $D/bench_under_ncu.log 2>&1
tail -3 $D/launches.csv
FQEB_OZAKI_PROF=0 timeout -s KILL 500 ncu --set full --clock-control none --import-source on \
  -k regex:k_sigma_ozaki2 -s 4 -c 1 -o $D/ozaki2 -f python scripts/ozaki_prof.py 16 > $D/ncu_ozaki_log.txt 2>&1
tail -3 $D/ncu_ozaki_log.txt
FQEB_OZAKI_PROF=0 timeout -s KILL 500 ncu --set full --clock-control none --import-source on \
  -k regex:k_make_coeff -s 4 -c 1 -o $D/scatter -f python scripts/ozaki_prof.py 16 > $D/ncu_scatter_log.txt 2>&1
tail -3 $D/ncu_scatter_log.txt
ls -la $D/
