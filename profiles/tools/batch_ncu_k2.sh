#!/bin/bash
tag=${1:-k2v1}; st=${2:-global}; ropes=${3:-1184}
mkdir -p gpurun_out
export BRDAE_K2_STATE=$st
python bench.py --workload npendulum_n20_16k --systems $ropes --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$tag.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_$tag.log; exit 1; }
python -c "
import json; d=json.loads(open('gpurun_out/plain_$tag.log').read()); print('plain', round(d['value'],1), 'ropes/s', round(d['roofline']['kernel_ms'],1), 'ms')"
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --clock-control none --import-source on -k regex:k2_kernel -s 3 -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --workload npendulum_n20_16k --systems $ropes --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
