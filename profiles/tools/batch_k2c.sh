#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "npendulum" > gpurun_out/pytest_gpu_k2.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_k2.log
for st in global smem; do
  BRDAE_K2_STATE=$st timeout 600 python bench.py --workload npendulum_n20_16k --systems 8192 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/k2c_$st.log 2>gpurun_out/k2c_$st.err
  python -c "
import json; d=json.loads(open('gpurun_out/k2c_$st.log').read()); print('n20 x4096 $st', round(d['value'],1), 'ropes/s', round(d['roofline']['kernel_ms'],1), 'ms', d['config']['launch'])" || tail -3 gpurun_out/k2c_$st.err
done
