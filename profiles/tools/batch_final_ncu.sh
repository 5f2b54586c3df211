#!/bin/bash
# launch list of the default bench command (cold-cache, serialised per-launch times) + ncu --set full of the K1 kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_final.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1
echo "launch list rc=$?"
bash profiles/tools/batch_ncu_k1.sh v7
