#!/bin/bash
# full GPU suite with the new entry points + host C-ABI copy-thread sweep
D=gpurun_out/${1:-r2y}
mkdir -p $D
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $D/pytest_gpu.txt
for t in 2 4 8 16; do
  (FQEB_COPY_THREADS=$t timeout -s KILL 200 python scripts/cabi_host_time.py 16 3 >> $D/cabi_threads.txt 2>&1)
done
cat $D/cabi_threads.txt
