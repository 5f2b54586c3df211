#!/bin/bash
mkdir -p gpurun_out
for spec in "npendulum_n20_16k 16384" "npendulum_n50_16k 16384" "npendulum_n100_16k 4096"; do
  set -- $spec
  timeout 900 python bench.py --workload $1 --systems $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/k2d_$1.log 2>gpurun_out/k2d_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/k2d_$1.log').read()); print('$1 x$2', round(d['value'],1), 'ropes/s', round(d['roofline']['kernel_ms'],1), 'ms', round(d['roofline']['frac'],4), d['config']['launch'])" || tail -3 gpurun_out/k2d_$1.err
done
