#!/bin/bash
# batch 1 of this session: parity tests, then variants of the K1 pendulum kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_b1.log
bash profiles/tools/run_variants.sh default psel b128x2 psel_b128x2
