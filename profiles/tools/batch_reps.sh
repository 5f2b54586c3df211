#!/bin/bash
# sweep of the Newton-passes-per-cycle knob over the K1 workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b3.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_b3.log
for wl in pendulum3_1M_tend2 robertson_4M transistor_256k pendulum3_1M_tend10; do
  for r in 2 3 4 5 6; do
    BRDAE_NEWTON_REPS=$r python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/reps_${wl}_$r.log 2>gpurun_out/reps_${wl}_$r.err
    python -c "
import json; d=json.loads(open('gpurun_out/reps_${wl}_$r.log').read()); print('$wl reps=$r', round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/reps_${wl}_$r.err
  done
done
