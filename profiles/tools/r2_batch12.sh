#!/bin/bash
# Round 2, GPU call 12: K4 backward sweeps two nodes ahead (rod quantities of the previous node in registers); all ropes-per-warp geometries
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "npendulum or k4 or chain or dense" > gpurun_out/r2b12_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -2 gpurun_out/r2b12_pytest_chain.log
timeout -k 10 400 python profiles/tools/k4_variants.py bwd2 20:16384,50:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b12_k4.jsonl 2>gpurun_out/r2b12_k4.err; cat gpurun_out/r2b12_k4.jsonl
