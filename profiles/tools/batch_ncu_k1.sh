#!/bin/bash
# ncu --set full capture of the K1 pendulum kernel (262,144 systems), after the same command ran clean
tag=${1:-v5}
mkdir -p gpurun_out
python bench.py --systems 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$tag.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_$tag.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 3 -c 1 -f -o gpurun_out/prof_k1_$tag \
  python bench.py --systems 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/
