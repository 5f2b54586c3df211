#!/bin/bash
# Round 2, GPU call 18: CTA shapes of the Robertson and transistor kernels (two half-size CTAs per SM, as the pendulum kernel has)
mkdir -p gpurun_out; rm -f gpurun_out/r2b18.jsonl
for lib in "" s2 s3; do
  if [ -n "$lib" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$lib.so; tag=$lib; else unset BRDAE_LIB; tag=default; fi
  for wl in robertson transistor; do
    timeout -k 10 300 python profiles/tools/k1_variants.py $tag $wl "X=1" >> gpurun_out/r2b18.jsonl 2>>gpurun_out/r2b18.err
  done
done
cat gpurun_out/r2b18.jsonl; tail -2 gpurun_out/r2b18.err
