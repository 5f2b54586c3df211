#!/bin/bash
# Round 2, GPU call 10: K1 scheduling knobs after the pattern-restricted factors (Newton passes per cycle, CTA shape)
mkdir -p gpurun_out
timeout -k 10 300 python profiles/tools/k1_variants.py default pendulum "BRDAE_NEWTON_REPS=4;BRDAE_NEWTON_REPS=3;BRDAE_NEWTON_REPS=5" > gpurun_out/r2b10_k1_pend.jsonl 2>gpurun_out/r2b10.err; cat gpurun_out/r2b10_k1_pend.jsonl
BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_b256.so timeout -k 10 300 python profiles/tools/k1_variants.py b256 pendulum "BRDAE_NEWTON_REPS=4;BRDAE_NEWTON_REPS=3" > gpurun_out/r2b10_k1_pend_b256.jsonl 2>>gpurun_out/r2b10.err; cat gpurun_out/r2b10_k1_pend_b256.jsonl
timeout -k 10 300 python profiles/tools/k1_variants.py default robertson "BRDAE_NEWTON_REPS=3;BRDAE_NEWTON_REPS=2;BRDAE_NEWTON_REPS=4" > gpurun_out/r2b10_k1_rob.jsonl 2>>gpurun_out/r2b10.err; cat gpurun_out/r2b10_k1_rob.jsonl
timeout -k 10 300 python profiles/tools/k1_variants.py default transistor "BRDAE_NEWTON_REPS=3;BRDAE_NEWTON_REPS=2;BRDAE_NEWTON_REPS=4" > gpurun_out/r2b10_k1_tra.jsonl 2>>gpurun_out/r2b10.err; cat gpurun_out/r2b10_k1_tra.jsonl
tail -3 gpurun_out/r2b10.err
