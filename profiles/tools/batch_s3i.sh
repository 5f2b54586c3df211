#!/bin/bash
# final evidence of round 1 for the K1 with fast paths (v8): the default bench line (with e2e and both CPU comparators),
# the launch list of the default bench command, and one ncu --set full capture of the pendulum kernel
mkdir -p gpurun_out
python bench.py > gpurun_out/final2_default.log 2> gpurun_out/final2_default.err; echo "default rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_v8.log 2>&1
echo "launch list rc=$?"
bash profiles/tools/batch_ncu_k1.sh v8 | tail -3
