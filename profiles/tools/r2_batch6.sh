#!/bin/bash
# Round 2, GPU call 6: K4 with the node stream staged through shared memory by TMA bulk copies -- chain parity tests first
# (bit-identical to the C oracle), then timings against the unstaged build.
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "npendulum or k4 or chain" > gpurun_out/r2b6_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -4 gpurun_out/r2b6_pytest_chain.log
S="BRDAE_K4_VOTE=1;BRDAE_K4_VOTE=1 BRDAE_K4_RPW=32;BRDAE_K4_VOTE=1 BRDAE_K4_RPW=8;BRDAE_K4_VOTE=0"
timeout -k 10 300 python profiles/tools/k4_variants.py staged 20:16384 "$S" > gpurun_out/r2b6_k4_staged.jsonl 2>gpurun_out/r2b6_k4_staged.err; cat gpurun_out/r2b6_k4_staged.jsonl; tail -3 gpurun_out/r2b6_k4_staged.err
BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_unstaged.so timeout -k 10 300 python profiles/tools/k4_variants.py unstaged 20:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b6_k4_unstaged.jsonl 2>gpurun_out/r2b6_k4_unstaged.err; cat gpurun_out/r2b6_k4_unstaged.jsonl
timeout -k 10 400 python profiles/tools/k4_variants.py staged 50:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b6_k4_big.jsonl 2>gpurun_out/r2b6_k4_big.err; cat gpurun_out/r2b6_k4_big.jsonl
BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_unstaged.so timeout -k 10 400 python profiles/tools/k4_variants.py unstaged 50:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b6_k4_big_unstaged.jsonl 2>gpurun_out/r2b6_k4_big_unstaged.err; cat gpurun_out/r2b6_k4_big_unstaged.jsonl
