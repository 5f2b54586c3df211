#!/bin/bash
mkdir -p gpurun_out
for st in smem global; do
for wl in npendulum_n20_16k; do
  BRDAE_K2_STATE=$st timeout 600 python bench.py --workload $wl --systems 8192 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/k2_${wl}_$st.log 2>gpurun_out/k2_${wl}_$st.err
  python -c "
import json; d=json.loads(open('gpurun_out/k2_${wl}_$st.log').read()); print('$wl $st', round(d['value'],1), 'ropes/s', round(d['roofline']['kernel_ms'],1), 'ms', round(d['roofline']['frac'],4), d['config']['launch'])" || tail -3 gpurun_out/k2_${wl}_$st.err
done; done
