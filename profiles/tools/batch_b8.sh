#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b8.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_b8.log
timeout 300 python profiles/tools/e2e_probe2.py 2>&1 | tail -6
