#!/bin/bash
# Round 2, GPU call 13: Newton node loops unrolled by two (register moves between visits) against the rolled default
mkdir -p gpurun_out
timeout -k 10 300 python profiles/tools/k4_variants.py rolled 20:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b13_k4_rolled.jsonl 2>gpurun_out/r2b13.err; cat gpurun_out/r2b13_k4_rolled.jsonl
BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_u2.so timeout -k 10 300 python profiles/tools/k4_variants.py u2 20:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b13_k4_u2.jsonl 2>>gpurun_out/r2b13.err; cat gpurun_out/r2b13_k4_u2.jsonl
