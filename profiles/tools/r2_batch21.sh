#!/bin/bash
# Round 2, GPU call 21: Newton vote threshold on the long chains
mkdir -p gpurun_out
timeout -k 10 600 python profiles/tools/k4_variants.py vote 50:16384,100:16384 "BRDAE_K4_VOTE=1;BRDAE_K4_VOTE=2;BRDAE_K4_VOTE=4;BRDAE_K4_VOTE=8;BRDAE_K4_VOTE=1 BRDAE_NEWTON_REPS=5" > gpurun_out/r2b21.jsonl 2>gpurun_out/r2b21.err; cat gpurun_out/r2b21.jsonl
