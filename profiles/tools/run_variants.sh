#!/bin/bash
# Tuning helper: bench every dae_scipy_b200/libbrdae_<tag>.so given on the command line ("" = default build).
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" != "default" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$v.so; else unset BRDAE_LIB; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$v.log 2>gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/var_$v.log').read()); print('$v', round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4), d['config']['launch'])"
done
