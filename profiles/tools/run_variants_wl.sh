#!/bin/bash
# usage: run_variants_wl.sh <workload> <extra bench args...> -- <variant tags...>   ("default" = libbrdae.so)
wl=$1; shift; extra=(); while [ "$1" != "--" ]; do extra+=("$1"); shift; done; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" != "default" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$v.so; else unset BRDAE_LIB; fi
  python bench.py --workload $wl "${extra[@]}" --no-cpu-baseline --no-e2e > gpurun_out/var_$v.log 2>gpurun_out/var_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/var_$v.log').read()); print('$v', round(d['value'],1), round(d['roofline']['kernel_ms'],1), round(d['roofline']['frac'],4), d['config']['launch'])" || tail -3 gpurun_out/var_$v.err
done
