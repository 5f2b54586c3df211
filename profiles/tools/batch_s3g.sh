#!/bin/bash
# A/B of two more K1 fast paths (same source, one switched off per variant library): the first-iteration convergence
# test (-DBRDAE_K1_FIRST_ITER_FASTPATH=0) and the optimistic factorisation without row exchanges (-DBRDAE_K1_STATIC_LU=0)
mkdir -p gpurun_out
for wl in pendulum3_1M_tend2 robertson_4M transistor_256k; do
  for v in default nofirst nostatic; do
    if [ $v = default ]; then unset BRDAE_LIB; else export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$v.so; fi
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab2_${wl}_$v.log 2> gpurun_out/ab2_${wl}_$v.err
    python -c "
import json; d=json.loads(open('gpurun_out/ab2_${wl}_$v.log').read()); print('$wl $v', round(d['value']), round(d['roofline']['kernel_ms'],2), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab2_${wl}_$v.err
  done
done
unset BRDAE_LIB
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
