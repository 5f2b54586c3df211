#!/bin/bash
# Round 2, GPU call 8: K4 with the first output point produced by the accept sweep: parity tests, timings of the three chain workloads
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "npendulum or k4 or chain or dense" > gpurun_out/r2b8_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -2 gpurun_out/r2b8_pytest_chain.log
timeout -k 10 400 python profiles/tools/k4_variants.py teval 20:16384,50:16384,100:16384,20:65536 "BRDAE_K4_VOTE=1" > gpurun_out/r2b8_k4.jsonl 2>gpurun_out/r2b8_k4.err; cat gpurun_out/r2b8_k4.jsonl
