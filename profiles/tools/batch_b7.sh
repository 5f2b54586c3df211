#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b7.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_b7.log
CHUNKS=131072 timeout 300 python profiles/tools/e2e_probe.py
BRDAE_ROWS_CHUNKED=1 CHUNKS=524288 timeout 300 python profiles/tools/e2e_probe.py
