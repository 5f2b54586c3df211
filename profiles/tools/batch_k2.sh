#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "npendulum or dense_output" > gpurun_out/pytest_gpu_k2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_k2.log
for wl in npendulum_n20_16k npendulum_n50_16k; do
  timeout 600 python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/k2_$wl.log 2>gpurun_out/k2_$wl.err
  python -c "
import json; d=json.loads(open('gpurun_out/k2_$wl.log').read()); print('$wl', round(d['value'],1), 'ropes/s', round(d['roofline']['kernel_ms'],1), 'ms', round(d['roofline']['frac'],4), d['config']['launch'])" || tail -3 gpurun_out/k2_$wl.err
done
