#!/bin/bash
# round-1 final K1 (fast paths gated by per-problem traits): device-timed lines of the K1 workloads and the GPU suite
mkdir -p gpurun_out
for wl in pendulum3_1M_tend2 pendulum3_1M_tend10 robertson_4M transistor_256k; do
  python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/fin2_${wl}.log 2> gpurun_out/fin2_${wl}.err
  python -c "
import json; d=json.loads(open('gpurun_out/fin2_${wl}.log').read()); print('$wl', round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/fin2_${wl}.err
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
