#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b2.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_b2.log
bash profiles/tools/run_variants.sh "$@"
