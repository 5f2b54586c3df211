#!/bin/bash
# Round 2, GPU call 20: pendulum kernel, per-warp cycle (no CTA-wide barrier) and four 64-thread CTAs per SM, re-tested on the smaller round-2 code
mkdir -p gpurun_out; rm -f gpurun_out/r2b20.jsonl
for lib in "" cw0 b64; do
  if [ -n "$lib" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$lib.so; tag=$lib; else unset BRDAE_LIB; tag=default; fi
  timeout -k 10 300 python profiles/tools/k1_variants.py $tag pendulum "X=1" >> gpurun_out/r2b20.jsonl 2>>gpurun_out/r2b20.err
done
cat gpurun_out/r2b20.jsonl; tail -2 gpurun_out/r2b20.err
