#!/bin/bash
# Round 2, GPU call 14: tableau constants from constant memory (default) against 64-bit immediates (variant imm)
mkdir -p gpurun_out
for lib in "" imm; do
  if [ -n "$lib" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$lib.so; tag=$lib; else unset BRDAE_LIB; tag=cbank; fi
  for wl in pendulum robertson transistor; do
    timeout -k 10 300 python profiles/tools/k1_variants.py $tag $wl "X=1" >> gpurun_out/r2b14.jsonl 2>>gpurun_out/r2b14.err
  done
  timeout -k 10 300 python profiles/tools/k4_variants.py $tag 20:16384 "BRDAE_K4_VOTE=1" >> gpurun_out/r2b14.jsonl 2>>gpurun_out/r2b14.err
done
cat gpurun_out/r2b14.jsonl
unset BRDAE_LIB
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_identical_to_c_oracle or npendulum" > gpurun_out/r2b14_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2b14_pytest.log
