#!/bin/bash
# bench every single-GPU workload once (device-timed value, kernel ms, roofline fraction)
mkdir -p gpurun_out
for w in pendulum3_1M_tend2 pendulum3_1M_tend10 robertson_4M transistor_256k; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/wl_$w.log 2>gpurun_out/wl_$w.err
  python -c "
import json; d=json.loads(open('gpurun_out/wl_$w.log').read()); print('$w', round(d['value']), 'traj/s', round(d['roofline']['kernel_ms'],2), 'ms', 'fp64 frac', round(d['roofline']['frac'],4), d['config']['launch'], d['status_histogram'], 'mean steps', round(d['mean_accepted_steps'],1))" || tail -3 gpurun_out/wl_$w.err
done
